import random

from masp_b200 import synthetic as syn

R = syn.R_INT
ib = syn.ints_to_bytes


def rand_scalars(n, seed, kinds="mixed"):
    rnd = random.Random(seed)
    if kinds == "uniform":
        return [rnd.randrange(R) for _ in range(n)]
    pool = [lambda: 0, lambda: 1, lambda: rnd.randrange(R), lambda: rnd.randrange(R), lambda: R - 1,
            lambda: rnd.randrange(1 << 16), lambda: 1 << rnd.randrange(255), lambda: 2]
    return [rnd.choice(pool)() for _ in range(n)]


def assignment(pv, w):
    return pv.ProvingAssignment(w["a"], w["b"], w["c"], w["inputs"], w["aux"])


def oracle_proofs(co, key_bytes, shape, dens, ws):
    P = co.Params(key_bytes, shape.n_aux, *dens)
    return [P.prove(shape.rows, w["a"], w["b"], w["c"], w["inputs"], w["aux"], w["r"], w["s"]) for w in ws]
