#!/usr/bin/env python
"""Extracts the reference's own golden values for the key / note layer above the circuits.

  * the five fixed generators and the six Pedersen-hash generators, as the (u, v) limbs written
    in masp_primitives/src/constants.rs:50-251 -- the reference re-derives each of them with
    find_group_hash in its tests (constants.rs:305-375), which pins group_hash, the BLAKE2s
    personalizations, point decompression and cofactor clearing;
  * the note-commitment vectors of masp_primitives/src/test_vectors/note_encryption.rs
    (default_d, default_pk_d, v, rcm, cmu), which the reference checks with
    `to.create_note(asset_type, tv.v, Rseed::BeforeZip212(rcm)).cmu() == cmu` under the asset
    identifier b"testtesttesttesttesttesttesttest" (sapling/note_encryption.rs:1324, 1357-1361).

The reference tree is only readable in the build container, so the parsed data is committed as
tests/golden/sapling_vectors.json.

    python tests/golden/make_sapling_vectors.py [/root/reference]
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
prim = os.path.join(ref, "masp_primitives/src")


def limbs(block):
    """[u, v] of every `from_u64s_le(&[l0, l1, l2, l3])` pair in a block of source."""
    vals = []
    for m in re.finditer(r"from_u64s_le\(&\[(.*?)\]\)", block, re.S):
        ls = [int(x.replace("_", ""), 16) for x in re.findall(r"0x([0-9a-f_]+)", m.group(1))]
        assert len(ls) == 4
        vals.append(sum(l << (64 * i) for i, l in enumerate(ls)))
    assert len(vals) % 2 == 0
    return [["%064x" % vals[i], "%064x" % vals[i + 1]] for i in range(0, len(vals), 2)]


src = open(os.path.join(prim, "constants.rs")).read()
gens = {}
for name in ("proof_generation_key_generator", "note_commitment_randomness_generator", "nullifier_position_generator",
             "value_commitment_randomness_generator", "spending_key_generator", "pedersen_hash_generators"):
    body = re.search(r"pub fn %s\(\) -> .*?\n\}" % name, src, re.S).group(0)
    pts = limbs(body)
    gens[name] = pts if name == "pedersen_hash_generators" else pts[0]
assert len(gens["pedersen_hash_generators"]) == 6

src = open(os.path.join(prim, "test_vectors/note_encryption.rs")).read()
src = src[src.index("pub fn make_test_vectors"):]  # skip the struct definition
notes = []
for m in re.finditer(r"TestVector \{(.*?)\n        \}", src, re.S):
    body = m.group(1)

    def field(name):
        arr = re.search(r"\b%s: \[(.*?)\]" % name, body, re.S).group(1)
        return bytes(int(x, 16) for x in re.findall(r"0x([0-9a-f]{2})", arr)).hex()

    notes.append({"ivk": field("ivk"), "default_d": field("default_d"), "default_pk_d": field("default_pk_d"),
                  "v": int(re.search(r"\bv: (\d+)", body).group(1)), "rcm": field("rcm"), "cmu": field("cmu")})
assert notes and all(len(n["default_d"]) == 22 and len(n["cmu"]) == 64 for n in notes)

# Merkle tree: HEX_EMPTY_ROOTS (merkle_tree.rs:912-946, checked by empty_root_test_vectors :1055-1063) and the
# commitments / roots of test_sapling_tree (:1091-1135): root[i] is the depth-32 root after i + 1 appends.
src = open(os.path.join(prim, "merkle_tree.rs")).read()
blk = src[src.index("const HEX_EMPTY_ROOTS"):]
empty_roots = re.findall(r'"([0-9a-f]{64})"', blk[:blk.index("];")])
assert len(empty_roots) == 33
blk = src[src.index("fn test_sapling_tree"):]
c0 = blk.index("let commitments = [")
commitments = re.findall(r'"([0-9a-f]{64})"', blk[c0:blk.index("];", c0)])
r0 = blk.index("let roots = [")
roots = re.findall(r'"([0-9a-f]{64})"', blk[r0:blk.index("];", r0)])
assert len(commitments) == 16 and len(roots) == 16

# ZIP 32 vectors (zip32/sapling.rs:1372-2135): ask, nsk -> ak, nk -> ivk (and the internal branch), plus the
# diversifiers the reference found valid (d0, d1, d2, dmax).  They pin ak = ask * G_spend, nk = nsk * G_proof and
# CRH^ivk (ViewingKey::ivk, sapling.rs:338-355) on values the reference asserts at zip32/sapling.rs:2074-2106.
src = open(os.path.join(prim, "zip32/sapling.rs")).read()
src = src[src.index("let test_vectors = vec!["):]
src = src[:src.index("assert_eq!(test_vectors.len()")]
keys = []
for m in re.finditer(r"TestVector \{(.*?)\n            \},", src, re.S):
    body = m.group(1)

    def opt(name):
        mm = re.search(r"\b%s: (None|(?:Some\()?\[(.*?)\])" % name, body, re.S)
        if mm is None or mm.group(1) == "None":
            return None
        return bytes(int(x, 16) for x in re.findall(r"0x([0-9a-f]{2})", mm.group(2))).hex()

    keys.append({k: opt(k) for k in ("ask", "nsk", "ak", "nk", "ivk", "d0", "d1", "d2", "dmax", "internal_nsk",
                                     "internal_nk", "internal_ivk")})
assert len(keys) == 5 and all(len(k["ak"]) == 64 and len(k["ivk"]) == 64 for k in keys)

out = {"asset_identifier": b"testtesttesttesttesttesttesttest".hex(), "generators": gens, "note_commitments": notes,
       "zip32_keys": keys, "empty_roots": empty_roots, "tree_commitments": commitments, "tree_roots": roots}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sapling_vectors.json")
json.dump(out, open(dst, "w"), indent=0)
print(len(notes), "note vectors, 11 generators, 33 empty roots, 16 tree roots, %d key vectors ->" % len(keys), dst)
