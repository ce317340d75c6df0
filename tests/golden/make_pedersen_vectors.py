#!/usr/bin/env python
"""Extracts the reference's own Pedersen-hash golden vectors into a JSON fixture.

Source: masp_primitives/src/test_vectors/pedersen_hash_vectors.rs (37 vectors, originally from
zcash-test-vectors sapling_pedersen.py), checked by the reference's test
masp_primitives/src/sapling/pedersen_hash.rs:131-152.  The reference tree is only readable in the
build container, so the parsed data is committed as tests/golden/pedersen_hash_vectors.json.

    python tests/golden/make_pedersen_vectors.py [/root/reference]
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(ref, "masp_primitives/src/test_vectors/pedersen_hash_vectors.rs")).read()
out = []
for m in re.finditer(r"TestVector \{(.*?)\n        \}", src, re.S):
    body = m.group(1)
    pers = re.search(r"personalization: Personalization::(\w+)(?:\((\d+)\))?", body)
    bits = [int(x) for x in re.findall(r"\b[01]\b", re.search(r"input_bits: vec!\[(.*?)\]", body, re.S).group(1))]
    u = re.search(r'hash_u: "Scalar\(0x([0-9a-f]{64})\)"', body).group(1)
    v = re.search(r'hash_v: "Scalar\(0x([0-9a-f]{64})\)"', body).group(1)
    out.append({"personalization": pers.group(1) if pers.group(2) is None else "MerkleTree(%s)" % pers.group(2),
                "input_bits": "".join(map(str, bits)), "hash_u": u, "hash_v": v})
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pedersen_hash_vectors.json")
json.dump(out, open(dst, "w"), indent=0)
print(len(out), "vectors ->", dst)
