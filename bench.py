#!/usr/bin/env python
"""Headline benchmark: Groth16 proofs/s for the MASP Spend circuit shape
(BASELINE.json configs[1]: Spend, ~2^17 constraints, batch = 256 proofs per
GPU, synthetic witnesses, explicit r/s), plus the other BASELINE configs in the
same JSON line (`configs`: Output, Convert x1024, the mixed 4096-tx batch, the
standalone G1 MSM sweep).

  python bench.py --gpus 1 --steps K --warmup W            # this framework, one GPU
  torchrun --nproc-per-node N bench.py --gpus N ...        # one process per GPU (the driver's launch)
  python bench.py --gpus N ...                              # ONE process driving N GPUs through the C ABI
  python bench.py --impl reference --gpus 1 --steps K ...  # restated bellperson CPU prover on the host cores

A step is one pass of the hot path over one batch: for every proof the H
pipeline (NTTs), the four bucket MSMs (H+L, A, B1 over G1; B2 over G2), assembly
and 192-byte encoding.  `value` is measured with the batch already resident
in HBM; `e2e` goes through mb200_prove_submit / _wait with pinned HOST buffers
(host->device copies of every witness and the device->host read of the
proofs inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from masp_b200 import synthetic as syn  # noqa: E402


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples),
                "reasons": sorted(reasons)}


def make_batch(shape, batch, first_index, torch, pv, device_copy=True):
    """Synthetic witnesses for proofs [first_index, first_index + batch) in
    pinned host memory (SURVEY §8d distribution); c = a * b by the device
    kernel behind mb200_fr_mul_device."""
    rows, n_aux, n_in = shape.rows, shape.n_aux, shape.n_inputs
    pin = lambda n: torch.empty(n, dtype=torch.uint8).pin_memory()
    host = {"a": pin(batch * rows * 32), "b": pin(batch * rows * 32), "c": pin(batch * rows * 32),
            "aux": pin(batch * n_aux * 32), "inputs": pin(batch * n_in * 32), "r": pin(batch * 32), "s": pin(batch * 32)}
    view = {k: v.numpy().view("<u8") for k, v in host.items()}
    cls = shape.aux_classes()
    is_bool = (cls & 4) != 0
    for i in range(batch):
        base = syn.STREAM_WIT_A + 8 * (first_index + i)
        a = syn.fr_uniform(syn.MASTER_SEED, base + 0, rows)
        b = syn.fr_uniform(syn.MASTER_SEED, base + 1, rows)
        inputs = syn.fr_uniform(syn.MASTER_SEED, base + 2, n_in)
        inputs[0] = (1, 0, 0, 0)
        a[shape.n_constraints:] = inputs
        b[shape.n_constraints:] = 0
        aux = syn.fr_uniform(syn.MASTER_SEED, base + 3, n_aux)
        bits = syn.fr_bits(syn.MASTER_SEED, base + 4, n_aux)
        aux[is_bool] = bits[is_bool]
        rs = syn.fr_uniform(syn.MASTER_SEED, base + 5, 2)
        view["a"][i * rows * 4:(i + 1) * rows * 4] = a.reshape(-1)
        view["b"][i * rows * 4:(i + 1) * rows * 4] = b.reshape(-1)
        view["aux"][i * n_aux * 4:(i + 1) * n_aux * 4] = aux.reshape(-1)
        view["inputs"][i * n_in * 4:(i + 1) * n_in * 4] = inputs.reshape(-1)
        view["r"][i * 4:(i + 1) * 4] = rs[0]
        view["s"][i * 4:(i + 1) * 4] = rs[1]
    # c = a * b on the device, in slabs (the pool of one circuit may be a few GB)
    slab = max(1, min(batch, (1 << 28) // (rows * 32)))
    for lo in range(0, batch, slab):
        hi = min(batch, lo + slab)
        sl = slice(lo * rows * 32, hi * rows * 32)
        da, db = host["a"][sl].cuda(non_blocking=True), host["b"][sl].cuda(non_blocking=True)
        dc = torch.empty_like(da)
        torch.cuda.synchronize()
        pv.fr_mul_device(da, db, (hi - lo) * rows, dc)
        host["c"][sl].copy_(dc)
        torch.cuda.synchronize()
    dev = None
    if device_copy:
        dev = {k: v.cuda(non_blocking=True) for k, v in host.items()}
        torch.cuda.synchronize()
    return host, dev


class Workload:
    """One circuit's key on the device(s) and a pool of synthetic witnesses in pinned host memory."""

    def __init__(self, pv, torch, shape, pool, first_index, device_copy=False):
        self.pv, self.torch, self.shape, self.pool = pv, torch, shape, pool
        t0 = time.perf_counter()
        self.key = pv.params_synthesize(shape)
        self.params = pv.Parameters.read(self.key, shape.densities())
        self.t_key = time.perf_counter() - t0
        self.first_index = first_index
        self.host, self.dev = make_batch(shape, pool, first_index, torch, pv, device_copy)
        self.t_setup = time.perf_counter() - t0

    def bytes_per_proof_h2d(self):
        sh = self.shape
        return 32 * (3 * sh.rows + sh.n_aux + sh.n_inputs + 2)

    def submit(self, lo, n, out, device=False):
        """Proofs [lo, lo + n) of the pool -> ticket; `out` a pinned uint8 tensor of n * 192 bytes."""
        sh, src = self.shape, (self.dev if device else self.host)
        per = {"a": sh.rows, "b": sh.rows, "c": sh.rows, "inputs": sh.n_inputs, "aux": sh.n_aux, "r": 1, "s": 1}
        v = {k: src[k][lo * per[k] * 32:(lo + n) * per[k] * 32] for k in per}
        return self.pv.prove_submit(self.params, n, sh.rows, v["a"], v["b"], v["c"], v["inputs"], v["aux"], v["r"],
                                    v["s"], out, device=device)

    def oracle_check(self, co, proofs, indices):
        """Byte comparison of the GPU's proofs (pool order, 192 bytes each) with the CPU oracle."""
        sh = self.shape
        P = co.Params(self.key, sh.n_aux, *sh.densities())
        g = lambda k, per, i: bytes(self.host[k].numpy()[i * per * 32:(i + 1) * per * 32])
        same, secs = 0, []
        for i in indices:
            t0 = time.perf_counter()
            want = P.prove(sh.rows, g("a", sh.rows, i), g("b", sh.rows, i), g("c", sh.rows, i), g("inputs", sh.n_inputs, i),
                           g("aux", sh.n_aux, i), g("r", 1, i), g("s", 1, i))
            secs.append(time.perf_counter() - t0)
            same += want == proofs[192 * i:192 * (i + 1)]
        return same, secs


def run_queue(pv, torch, items, in_flight):
    """A longest-first work queue: items = [(workload, lo, n, out)], submitted in order with at most
    `in_flight` batches outstanding (the tail of one batch runs under the heads of the next ones)."""
    tickets = []
    for wl, lo, n, out in items:
        tickets.append(wl.submit(lo, n, out))
        if len(tickets) >= in_flight:
            pv.prove_wait(tickets.pop(0))
    while tickets:
        pv.prove_wait(tickets.pop(0))


def circuit_path(pv, key, shape, n, rounds, seed=20261017, torch=None, threads=0, sync=None):
    """The drop-in call a TxProver makes, measured: real Spend witnesses ->
    product-side witness generation on the host cores (mb200_circuit_synthesize)
    -> only inputs + aux cross PCIe -> rows on the device (r1cs_eval) -> proof.
    Host synthesis of batch k+1 overlaps the device work of batch k."""
    import queue
    import random
    from masp_b200 import circuits as C
    if shape.name != "spend":
        return None
    t0 = time.perf_counter()
    circ = C.Circuit(C.SPEND)
    t_record = time.perf_counter() - t0
    assert (circ.n_constraints, circ.n_inputs, circ.hash()) == C.PINS[C.SPEND]
    params = pv.Parameters.read(key, circ.densities()).bind_circuit(circ)
    rnd = random.Random(seed)
    js = lambda: rnd.randrange(C.JUBJUB_ORDER)

    def instances(k):
        out = []
        for _ in range(k):
            path = [(rnd.randrange(syn.R_INT), bool(rnd.getrandbits(1))) for _ in range(C.TREE_DEPTH)]
            inst = C.Spend(C.ValueCommitmentOpening(C.PROOF_GENERATION_KEY_GENERATOR, rnd.getrandbits(64), js()),
                           C.SPENDING_KEY_GENERATOR, js(), C.PROOF_GENERATION_KEY_GENERATOR, js(), js(), path, 0)
            out.append(inst)
        return out
    batches = [instances(n) for _ in range(rounds)]
    for inst in batches[0][:2]:          # satisfied witnesses where it is cheap to make them so
        inst.anchor = circ.root(inst)
    packed = [[i.pack() for i in b] for b in batches]
    to_b = lambda vals: b"".join(int(v).to_bytes(32, "little") for v in vals)
    r_b, s_b = to_b(rnd.randrange(syn.R_INT) for _ in range(n)), to_b(rnd.randrange(syn.R_INT) for _ in range(n))
    outs = [np.empty(192 * n, dtype=np.uint8) for _ in range(rounds)]

    # witnesses are generated straight into pinned memory (a ring of buffer pairs), so the
    # host-to-device copies of a submission are asynchronous
    def pinned(nbytes):
        if torch is None:
            return np.empty(nbytes, dtype=np.uint8)
        return torch.empty(nbytes, dtype=torch.uint8).pin_memory().numpy()
    ring = [(pinned(n * circ.n_inputs * 32), pinned(n * circ.n_aux * 32)) for _ in range(6)]
    # each stage alone
    t0 = time.perf_counter()
    inputs, aux = circ.synthesize(packed[0], out=ring[0], threads=threads)
    t_synth = time.perf_counter() - t0
    pv.prove_wait(pv.prove_submit_witness(params, n, inputs, aux, r_b, s_b, outs[0]))  # warm-up
    t0 = time.perf_counter()
    pv.prove_wait(pv.prove_submit_witness(params, n, inputs, aux, r_b, s_b, outs[0]))
    t_prove = time.perf_counter() - t0
    proofs = [outs[0][:192].tobytes()]

    def pipelined(in_flight):
        """A host thread synthesises batch after batch; this thread keeps up to `in_flight`
        batches on the device (submit / wait)."""
        qq, fr = queue.Queue(maxsize=1), queue.Queue()
        for b in ring[:in_flight + 2]:
            fr.put(b)

        def prod():
            for k in range(rounds):
                qq.put(circ.synthesize(packed[k], out=fr.get(), threads=threads))
        if sync:
            sync()
        t1 = time.perf_counter()
        thr = threading.Thread(target=prod)
        thr.start()
        tk, kp = [], []
        for k in range(rounds):
            i_k, a_k = qq.get()
            kp.append((i_k, a_k))
            tk.append(pv.prove_submit_witness(params, n, i_k, a_k, r_b, s_b, outs[k]))
            if len(tk) > in_flight:
                pv.prove_wait(tk.pop(0))
                fr.put(kp.pop(0))
        while tk:
            pv.prove_wait(tk.pop(0))
            fr.put(kp.pop(0))
        thr.join()
        dt = time.perf_counter() - t1
        return rounds * n / dt, dt
    plain, t_plain = pipelined(2)
    # the reference's spend_proof also runs verify_proof on the fresh proof (sapling/prover.rs:148):
    # same pipeline with the device self-check on (audit mode: this key is not a valid CRS, so the
    # verdicts are 'fail' by construction; the kernel and its cost are the same)
    pv.set_option("verify", 2)
    try:
        with_check, _ = pipelined(4)
    finally:
        pv.set_option("verify", 0)
    return {
        "what": "real Spend witnesses through mb200_circuit_synthesize + mb200_prove_batch_witness "
                "(what TxProver::spend_proof does per description, batched)",
        "proofs_per_s_pipelined": plain, "proofs_per_s_pipelined_with_self_check": with_check,
        "seconds_pipelined": t_plain, "batch": n, "rounds": rounds,
        "host_witness_per_s": n / t_synth, "host_threads": threads or os.cpu_count(),
        "device_proofs_per_s": n / t_prove,
        "h2d_bytes_per_proof": 32 * (circ.n_aux + circ.n_inputs + 2),
        "h2d_bytes_per_proof_with_rows": 32 * (3 * circ.rows + circ.n_aux + circ.n_inputs + 2),
        "circuit_record_s": round(t_record, 2), "cs_hash": circ.hash(),
        "matrix_nnz": [circ.nnz_a, circ.nnz_b, circ.nnz_c],
        "proof_bytes_sample": proofs[0][:8].hex(),
    }


def integer_roof(shape, c_hl, c_a, fpmul, proofs_per_s_per_gpu):
    """The bound that binds: Fp-multiplication equivalents per proof against the live
    multiplication rate of the chip.  Counts for this bench's witnesses (boolean aux are 0 or 1 with equal
    probability: 0 costs nothing, 1 one mixed addition; a full-width scalar one addition per window):
    G1 mixed addition = 10 products + 9 Montgomery reductions = 9.5 multiplications, G2 = 36 + 18 half
    units = 27, an Fr multiplication 128 / 276 of an Fp one; six transforms of m log2(m) / 2 butterflies."""
    nwin = lambda c: -(-254 // c)          # windows after folding s > (r - 1) / 2 (msm.cuh msm_nwin)
    sh = shape
    full_l = sh.n_aux - sh.n_bool
    full_a = sh.full_ab + sh.full_a
    full_b = sh.full_ab + sh.full_b
    bool_a = sh.bool_ab + sh.bool_a
    bool_b = sh.bool_ab + sh.bool_b
    g1 = (sh.h_len + full_l) * nwin(c_hl) + sh.n_bool / 2 + (full_a + full_b) * nwin(c_a) + (bool_a + bool_b) / 2
    g2 = full_b * nwin(c_a) + bool_b / 2
    fr = 6 * sh.m * sh.log_m / 2
    muls = 9.5 * g1 + 27.0 * g2 + fr * 128.0 / 276.0
    roof = fpmul / muls if muls else 0.0
    return {"fp_mul_equivalents_per_proof": muls, "g1_additions_per_proof": g1, "g2_additions_per_proof": g2,
            "fr_multiplications_per_proof": fr, "proofs_per_s_at_measured_fp_mul_rate": roof,
            "frac": (proofs_per_s_per_gpu / roof) if roof else None,
            "note": "bucket reductions, scalar multiplications of the assembly and to-affine conversions (< 4 %) not counted"}


def run_reference(args):
    """The reference's own CPU implementation of the path, restated
    (oracle/c: window-parallel Pippenger + radix-2 domain on all host cores;
    the Rust crates are not vendored and there is no cargo here).
    Spend is the headline; Output and Convert are timed beside it so that every
    circuit of BASELINE.json's metric has its CPU figure."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import c_oracle as co
    sample = args.ref_sample
    per_circuit = {}
    headline = None
    for name in ("spend", "output", "convert"):
        shape = syn.SHAPES[name]
        key = co.params_from_logs(syn.key_logs(shape))
        P = co.Params(key, shape.n_aux, *shape.densities())
        ws = [syn.witness(shape, i, co.fr_mul) for i in range(sample)]
        prove = lambda w: P.prove(shape.rows, w["a"], w["b"], w["c"], w["inputs"], w["aux"], w["r"], w["s"])
        steps, warm = (args.steps, args.warmup) if name == "spend" else (max(1, min(args.steps, 3)), 1)
        for _ in range(warm):
            prove(ws[0])
        t0 = time.perf_counter()
        for _ in range(steps):
            for w in ws:
                prove(w)
        dt = time.perf_counter() - t0
        per_circuit[name] = {"proofs_per_s": steps * sample / dt, "seconds_per_proof": dt / (steps * sample),
                             "proofs_timed": steps * sample}
        if name == "spend":
            headline = (steps * sample / dt, dt)
        del P, ws, key
    value, dt = headline
    cores = co.get_threads()
    line = {
        "impl": "reference", "metric": "spend_proofs_per_sec", "value": value, "unit": "proofs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (mod p 381-bit / mod r 255-bit)",
        "data": "synthetic",
        "config": {"workload": "configs[1]: Spend shape (rows 100645, m 2^17), explicit r/s; bounded sample of %d proofs per step" % sample,
                   "circuit": "spend", "batch_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": cores, "kind": "port",
                         "sample": "%d Spend-shaped proofs per step x %d steps, restated bellperson CPU prover (oracle/c)" % (sample, args.steps),
                         "circuits": per_circuit},
        "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def load_traffic():
    """DRAM bytes per launch of the four bucket-accumulation launches of one 64-proof Spend chunk,
    from the committed ncu capture (profiles/ncu_traffic.json names its source)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def msm_sweep(pv, torch, dist, rank, world, ndev, sizes, kinds, reps, check, emit=None):
    """BASELINE config 3: one G1 MSM of 2^k bases, range-split over the GPUs (world ranks under
    torch.distributed, or ndev devices inside this process); partials meet on a device (NCCL
    all-gather of 192-byte XYZZ points, or NVLink peer copies) and are added there."""
    from masp_b200 import sharding
    out = []
    local = torch.cuda.current_device()
    for log_n in sizes:
        n = 1 << log_n
        lo, hi = sharding.shard_range(n, rank, world)
        bases = pv.synth_points(syn.STREAM_MSM_BASE, lo, hi - lo, 1)
        if ndev > 1:
            gb = pv.SplitG1Bases(bases, n)
        else:
            gb = pv.G1Bases(bases, hi - lo)
        del bases
        for kind in kinds:
            sc_all = syn.msm_scalars(n, kind)
            sc = torch.from_numpy(np.ascontiguousarray(sc_all[lo:hi]).view(np.uint8).reshape(-1)).pin_memory()
            gathered = torch.zeros(world * 192, dtype=torch.uint8, device="cuda") if world > 1 else None
            mine = torch.zeros(192, dtype=torch.uint8, device="cuda")
            times, acc_us, result = [], [], None

            def one_call():
                if ndev > 1:
                    return gb.msm(sc)
                gb.msm_partial_into(sc, mine)
                if world > 1:
                    dist.all_gather_into_tensor(gathered, mine)   # NCCL, device buffers: K6's exchange
                    torch.cuda.current_stream().synchronize()     # the library adds on its own stream: order it after NCCL's
                    return pv.g1_sum_partials_device(gathered, world)
                return pv.g1_sum_partials_device(mine, 1)
            for rep in range(reps + 2):
                profiled = rep == reps + 1      # the last call times the accumulation kernels (event syncs inside)
                if profiled:
                    pv.set_option("profile", 1)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                result = one_call()
                dt = time.perf_counter() - t0
                if world > 1:
                    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dt = float(t.item())
                if profiled:
                    acc_us.append(pv.get_counter("acc_us"))
                    pv.set_option("profile", 0)
                elif rep:                       # the first call is warm-up
                    times.append(dt)
            if rank == 0:
                ok = None
                if check:
                    from oracle import c_oracle as co
                    logs = syn.fr_uniform(syn.MASTER_SEED, syn.STREAM_MSM_BASE, n)
                    dot = co.fr_dot(syn.limbs_to_bytes(sc_all), syn.limbs_to_bytes(logs), n)
                    ok = result == co.g1_gen_mul(dot.to_bytes(32, "little"), 1)
                    del logs
                best = min(times)
                acc = min(acc_us) * 1e-6 / max(1, ndev)   # the counter adds up the devices of this process
                if emit:
                    emit("msm 2^%d %s: %.2f ms" % (log_n, kind, 1e3 * best))
                out.append({"log_n": log_n, "scalars": kind, "n_gpus": world * ndev, "ms_total": 1e3 * best,
                            "ms_accumulate_kernel": 1e3 * acc,
                            "gbs_total": 128.0 * n / best / 1e9,
                            "gbs_accumulate_kernel_per_gpu": (128.0 * n / (world * ndev) / acc / 1e9) if acc else None,
                            "closed_form_ok": ok})
            del sc_all, sc
        del gb
    _ = local
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--circuit", default="spend", choices=list(syn.SHAPES))
    ap.add_argument("--batch", type=int, default=256, help="proofs per GPU per step")
    ap.add_argument("--chunk", type=int, default=0, help="proofs per in-flight chunk (0 = library default)")
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--ref-sample", type=int, default=2, help="--impl reference: proofs per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline budget on rank 0")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-circuit-path", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 2 / 4 and the Output line")
    ap.add_argument("--no-msm-sweep", action="store_true", help="skip BASELINE config 3")
    ap.add_argument("--msm-sizes", type=int, nargs="+", default=[16, 18, 20, 22, 24])
    ap.add_argument("--config-sample", type=int, default=64, help="proofs per config compared with the CPU oracle")
    ap.add_argument("--mixed-tx", type=int, default=512, help="config 4: transactions per GPU (4096 over 8 GPUs)")
    ap.add_argument("--verify", action="store_true",
                    help="run the Groth16 check on every proof inside the timed region (the reference's "
                         "verify_proof after create_random_proof, sapling/prover.rs:148)")
    ap.add_argument("--circuit-batch", type=int, default=64, help="real-witness leg: proofs per round")
    ap.add_argument("--circuit-rounds", type=int, default=12)
    args = ap.parse_args()
    shape = syn.SHAPES[args.circuit]

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # no torchrun and --gpus N > 1: ONE process opens N devices and the library shards every batch
    ndev = args.gpus if (world == 1 and "WORLD_SIZE" not in os.environ and args.gpus > 1) else 1
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import masp_b200.prover as pv
    pv.init(list(range(ndev)) if ndev > 1 else local)
    if args.chunk:
        pv.set_option("chunk", args.chunk)
    if args.streams:
        pv.set_option("streams", args.streams)
    if args.verify:
        # synthetic keys and witnesses do not verify: count, do not fail (same kernel, same cost)
        pv.set_option("verify", 2)
    n_gpus = world * ndev

    B = args.batch * ndev        # proofs per step handled by this process
    wl = Workload(pv, torch, shape, B, rank * B, device_copy=(ndev == 1))
    params, host, dev, key = wl.params, wl.host, wl.dev, wl.key
    rows = shape.rows
    out_ring = [torch.empty(B * 192, dtype=torch.uint8).pin_memory() for _ in range(2)]
    out_pinned = out_ring[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_sync(device):
        pv.prove_wait(wl.submit(0, B, out_pinned, device=device))
        return pv.get_counter("last_batch_us") * 1e-3

    def run_steps(k_steps, device):
        """K steps streamed through mb200_prove_submit / mb200_prove_wait with at
        most two batches in flight: the tail of step k overlaps the head of k+1."""
        tickets = []
        for k in range(k_steps):
            tickets.append(wl.submit(0, B, out_ring[k % 2], device=device))
            if len(tickets) > 1:
                pv.prove_wait(tickets.pop(0))
        while tickets:
            pv.prove_wait(tickets.pop(0))

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    resident = ndev == 1   # device-resident inputs live on one GPU: the single-process N-GPU mode has host inputs only
    dev_ms_sync = 0.0
    for _ in range(args.warmup):
        dev_ms_sync = step_sync(resident)
    proofs_dev = bytes(out_pinned.numpy())
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = pv.get_counter("launches")
    barrier()
    t0 = time.perf_counter()
    run_steps(args.steps, resident)
    barrier()
    wall = time.perf_counter() - t0
    launches = pv.get_counter("launches") - launches0
    wall = max_over_ranks(wall)
    dev_ms = max_over_ranks(dev_ms_sync) * args.steps
    assert bytes(out_ring[(args.steps - 1) % 2].numpy()) == proofs_dev, "streamed and synchronous calls disagree"

    # end to end: pinned host buffers in, proofs out, copies inside the timed region
    step_sync(False)
    barrier()
    t0 = time.perf_counter()
    run_steps(args.steps, False)
    barrier()
    wall_e2e = max_over_ranks(time.perf_counter() - t0)
    if rank == 0:
        sampler.stop_flag = True
    assert bytes(out_ring[(args.steps - 1) % 2].numpy()) == proofs_dev, "host-buffer and device-buffer paths disagree"

    # roofline of the dominant kernel (G1/G2 bucket accumulation), timed live with CUDA events per launch
    pv.set_option("profile", 1)
    step_sync(resident)
    acc_us, acc_n, acc_bytes = pv.get_counter("acc_us"), pv.get_counter("acc_launches"), pv.get_counter("acc_bytes")
    step_ms_profiled = pv.get_counter("last_batch_us") * 1e-3
    pv.set_option("profile", 0)
    fpmul = pv.bench_fpmul()

    # ---- the other BASELINE configs, same process, same streamed / pinned path ---------------------
    configs = {}
    cfg_checks = []   # (name, workload, proofs bytes, sample indices) for the CPU oracle on rank 0
    if not args.no_configs and shape.name == "spend":
        per_gpu = lambda k: k * ndev
        w_out = Workload(pv, torch, syn.OUTPUT, per_gpu(256), rank * per_gpu(256))
        w_cnv = Workload(pv, torch, syn.CONVERT, per_gpu(128), rank * per_gpu(128))

        def timed(items_fn, in_flight, passes):
            run_queue(pv, torch, items_fn(), in_flight)       # warm-up: every context grows its buffers
            barrier()
            t1 = time.perf_counter()
            for _ in range(passes):
                run_queue(pv, torch, items_fn(), in_flight)
            barrier()
            return max_over_ranks(time.perf_counter() - t1)

        # Output (no bench in the reference; the metric names it): 256 proofs per GPU per pass
        o_out = torch.empty(w_out.pool * 192, dtype=torch.uint8).pin_memory()
        dt = timed(lambda: [(w_out, 0, w_out.pool, o_out)], 2, 4)
        configs["output"] = {"proofs_per_s": n_gpus * 256 * 4 / dt, "per_gpu_batch": 256, "passes": 4, "seconds": dt,
                             "algorithmic_bytes_per_proof": syn.OUTPUT.algorithmic_bytes()}
        cfg_checks.append(("output", w_out, bytes(o_out.numpy())))
        # config 2: Convert x 1024 over 8 GPUs = 128 per GPU per pass (weak scaling: 128 per GPU at any N)
        o_cnv = torch.empty(w_cnv.pool * 192, dtype=torch.uint8).pin_memory()
        dt = timed(lambda: [(w_cnv, 0, w_cnv.pool, o_cnv)], 2, 8)
        configs["convert_1024"] = {"proofs_per_s": n_gpus * 128 * 8 / dt, "per_gpu_batch": 128, "passes": 8,
                                   "seconds": dt, "total_proofs_per_pass": n_gpus * 128,
                                   "algorithmic_bytes_per_proof": syn.CONVERT.algorithmic_bytes()}
        cfg_checks.append(("convert_1024", w_cnv, bytes(o_cnv.numpy())))
        # config 4: (2 Spend + 2 Output + 1 Convert) x 512 tx per GPU (4096 tx on 8 GPUs), ONE queue
        # ordered longest circuit first, cut into batches of `unit` proofs, <= 4 batches in flight.
        # Witness pools are cycled (256 Spend / 256 Output / 128 Convert distinct witnesses per GPU).
        tx = args.mixed_tx
        unit = 64 * ndev
        sink = {c: torch.empty(w.pool * 192, dtype=torch.uint8).pin_memory() for c, w in
                (("spend", wl), ("output", w_out), ("convert", w_cnv))}

        def mixed_items():
            items = []
            for w, total in ((wl, 2 * tx * ndev), (w_cnv, tx * ndev), (w_out, 2 * tx * ndev)):   # longest first
                done = 0
                while done < total:
                    lo = done % w.pool
                    n = min(unit, total - done, w.pool - lo)
                    o = sink[w.shape.name][lo * 192:(lo + n) * 192]
                    items.append((w, lo, n, o))
                    done += n
            return items
        items = mixed_items()
        run_queue(pv, torch, items[:8] + items[-8:] + items[len(items) // 2:len(items) // 2 + 4], 4)  # warm-up
        barrier()
        t1 = time.perf_counter()
        run_queue(pv, torch, items, 4)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t1)
        n_proofs = n_gpus * 5 * tx
        configs["mixed_4096tx"] = {"proofs_per_s": n_proofs / dt, "tx_per_s": n_gpus * tx / dt, "seconds": dt,
                                   "tx_total": n_gpus * tx, "proofs_total": n_proofs, "tx_per_gpu": tx,
                                   "queue": "longest first (Spend, Convert, Output), batches of %d proofs, <= 4 in flight" % unit,
                                   "witness_pools_per_gpu": {"spend": args.batch, "output": 256, "convert": 128}}
        assert bytes(sink["spend"].numpy()) == proofs_dev, "mixed queue: Spend proofs differ from the headline run"
        cfg_checks.append(("mixed_4096tx:output", w_out, bytes(sink["output"].numpy())))
        cfg_checks.append(("mixed_4096tx:convert", w_cnv, bytes(sink["convert"].numpy())))
        assert bytes(sink["output"].numpy()) == bytes(o_out.numpy()) and bytes(sink["convert"].numpy()) == bytes(o_cnv.numpy())

    sweep = None
    if not args.no_msm_sweep and shape.name == "spend":
        sweep = msm_sweep(pv, torch, dist, rank, world, ndev, args.msm_sizes, ["U", "W"], 2, check=True)

    # the drop-in call with REAL Spend witnesses, at any N: every rank generates its own witnesses on its
    # share of the host threads (one process: all of them) and proves them on its GPU(s)
    cpath = None
    if not args.no_circuit_path and shape.name == "spend":
        try:
            threads = max(1, (os.cpu_count() or 1) // world)
            cpath = circuit_path(pv, key, shape, args.circuit_batch * ndev, args.circuit_rounds, torch=torch,
                                 threads=threads, sync=barrier)
            if world > 1:
                t = max_over_ranks(cpath["seconds_pipelined"])
                cpath["proofs_per_s_pipelined_per_rank"] = cpath["proofs_per_s_pipelined"]
                cpath["proofs_per_s_pipelined"] = world * cpath["batch"] * cpath["rounds"] / t
                hw = torch.tensor([cpath["host_witness_per_s"]], dtype=torch.float64, device="cuda")
                dist.all_reduce(hw)
                cpath["host_witness_per_s"] = float(hw.item())
                cpath["host_threads_per_rank"] = threads
        except Exception as e:  # reported, never silently dropped
            cpath = {"error": repr(e)}
            if world > 1:
                raise

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    achieved = acc_bytes / (acc_us * 1e-6) / 1e9 if acc_us else 0.0
    ms_per_step = 1e3 * wall / args.steps
    value = world * B * args.steps / wall
    e2e_value = world * B * args.steps / wall_e2e
    h2d = B * wl.bytes_per_proof_h2d()
    tr = load_traffic() or {}
    traffic = tr.get("dram_bytes_per_launch_avg")
    alg_per_launch = acc_bytes / acc_n if acc_n else 0

    line = {
        "metric": "spend_proofs_per_sec" if shape.name == "spend" else shape.name + "_proofs_per_sec",
        "value": value, "unit": "proofs/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "device_ms_per_step_unpipelined": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (mod p 381-bit / mod r 255-bit)",
        "data": "synthetic",
        "config": {
            "workload": "configs[1]: %s shape (constraints %d, rows %d, m 2^%d), batch %d proofs per GPU, explicit r/s, "
                        "synthetic witnesses (%.1f%% boolean aux)" % (shape.name, shape.n_constraints, rows, shape.log_m,
                                                                    args.batch, 100.0 * shape.n_bool / shape.n_aux),
            "circuit": shape.name, "batch_per_gpu": args.batch,
            "parallelism": ("proof-sharded x%d, no collective; " % n_gpus) +
                           ("one process per GPU (torch.distributed)" if ndev == 1 else
                            "ONE process, %d devices behind the C ABI (mb200_init(ids, %d)); `value` uses pinned host "
                            "inputs like e2e: device-resident inputs live on a single GPU" % (ndev, ndev)),
            "self_check": "verify_proof on the device for every proof" if args.verify else "off",
            "l2": "inputs per step (%.2f GB) are larger than the 126 MB L2" % (h2d / 1e9),
            "pipelining": "steps are streamed (mb200_prove_submit / mb200_prove_wait, <= 2 batches in flight); "
                          "the timed region is bracketed by barrier + synchronize",
            "window_bits": {"h_l": params.window_hl, "a": params.window_a},
            "table_bytes_hbm": params.table_bytes, "algorithmic_bytes_per_proof": shape.algorithmic_bytes(),
            "setup_s": round(wl.t_setup, 2), "key_synth_and_load_s": round(wl.t_key, 2),
            # library knobs in effect (window sizes, chunking); empty = the shipped defaults
            "knobs": {k: v for k, v in sorted(os.environ.items()) if k.startswith("MB200_") and v not in ("", "0")},
        },
        "e2e": {"value": e2e_value, "unit": "proofs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * 192,
                "ms_per_step": 1e3 * wall_e2e / args.steps},
        "gpu_launches": int(launches),
        "self_check": ({"verified": int(pv.get_counter("verified")), "failed": int(pv.get_counter("verify_failed")),
                        "note": "synthetic keys / witnesses cannot verify; the count shows the kernel ran"}
                       if args.verify else None),
        "clocks": sampler.summary(),
        "roofline": {
            "bound": "hbm", "kernel": "msm_accumulate_g1/g2 (bucket accumulation, all four queries)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_over_algorithmic": (traffic / alg_per_launch) if (traffic and alg_per_launch) else None,
            "traffic_by_query": tr.get("by_query"), "traffic_source": tr.get("source"),
            "peak_source": peak_src, "launches_timed": int(acc_n),
            "algorithmic_bytes_per_launch": alg_per_launch,
            "avg_launch_ms": acc_us * 1e-3 / acc_n if acc_n else 0,
            "share_of_step": (acc_us * 1e-3 / ndev) / step_ms_profiled if step_ms_profiled else None,
            "note": "this path is bound by 32-bit integer multiply-add issue, not HBM (SURVEY.md §7/§8d): "
                    "see fp_mul_per_s for the integer-side figure",
            "fp_mul_per_s": fpmul,
            "integer_roof": integer_roof(shape, params.window_hl, params.window_a, fpmul, value / n_gpus),
            "whole_proof_hbm_frac": shape.algorithmic_bytes() * value / n_gpus / 1e9 / peak,
        },
    }
    if configs:
        line["configs"] = configs
    if sweep is not None:
        line.setdefault("configs", {})["msm_sweep"] = sweep

    if cpath is not None:
        cpath["n_gpus"] = n_gpus
        cpath["vs_synthetic_rows_e2e"] = (cpath["proofs_per_s_pipelined"] / e2e_value) if "proofs_per_s_pipelined" in cpath else None
        line["circuit_path"] = cpath

    if not args.no_cpu_baseline:
        from oracle import c_oracle as co
        # headline: the first proofs of the batch on all host threads, inside a time budget
        sh = shape
        P = co.Params(key, sh.n_aux, *sh.densities())
        g = lambda k, per, i: bytes(host[k].numpy()[i * per * 32:(i + 1) * per * 32])
        t0 = time.perf_counter()
        done, parity = 0, 0
        breakdown = np.zeros(8)
        while done < min(B, 16) and (done < 2 or time.perf_counter() - t0 < args.cpu_seconds):
            proof, tm = P.prove(rows, g("a", rows, done), g("b", rows, done), g("c", rows, done), g("inputs", sh.n_inputs, done),
                                g("aux", sh.n_aux, done), g("r", 1, done), g("s", 1, done), timings=True)
            breakdown += np.array(tm)
            parity += proof == proofs_dev[192 * done:192 * (done + 1)]
            done += 1
        dt = time.perf_counter() - t0
        del P
        line["cpu_baseline"] = {
            "value": done / dt, "unit": "proofs/s", "cores": co.get_threads(), "kind": "port",
            "sample": "first %d proofs of the same batch (%.1f s), restated bellperson CPU prover (oracle/c), "
                      "all host threads" % (done, dt),
            "seconds_per_proof": dt / done,
            "breakdown_s_per_proof": {"h_ntt_wall": breakdown[0] / done, "msm_wall": breakdown[1] / done,
                                      "assembly_wall": breakdown[2] / done},
            "gpu_proofs_byte_identical": "%d/%d" % (parity, done),
        }
        if parity != done:
            line["parity_error"] = "GPU proofs differ from the CPU oracle"
        # the configs: a sample of each compared byte for byte, and the CPU rate of that circuit
        for name, w, proofs in cfg_checks:
            if name.startswith("mixed"):
                k = args.config_sample // 4       # the mixed queue re-proves the pools checked below
            else:
                k = args.config_sample
            idx = [int(i) for i in np.linspace(0, w.pool - 1, num=min(k, w.pool)).astype(int)]
            same, secs = w.oracle_check(co, proofs, idx)
            tgt = line["configs"][name.split(":")[0]]
            tgt.setdefault("oracle_sample", {})[w.shape.name] = "%d/%d byte-identical" % (same, len(idx))
            if not name.startswith("mixed"):
                tgt["cpu_proofs_per_s"] = len(secs) / sum(secs)
                tgt["cpu_cores"] = co.get_threads()
            if same != len(idx):
                line["parity_error"] = "GPU proofs differ from the CPU oracle (%s)" % name
        if "mixed_4096tx" in line.get("configs", {}):
            line["configs"]["mixed_4096tx"]["oracle_sample"]["spend"] = "%d/%d byte-identical (headline batch = the Spend pool)" % (parity, done)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
