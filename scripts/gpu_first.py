"""First-contact GPU script: runs each kernel family once against the oracle and prints timings.
Run on the GPU box:  python scripts/gpu_first.py [stage ...]"""
import os, sys, time, random, struct, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cport, bn254 as bn, mimc7
from oracle import withdraw_circuit as wc
from oracle import groth16 as g16
import owshen_b200 as ob
from owshen_b200 import api

R, P = bn.R, bn.P
random.seed(1234)
ctx = ob.Context(0)
stages = sys.argv[1:] or ["imad", "field", "mimc", "ntt", "msm", "witness", "setup", "prove"]
ok = True


def stage(name):
    def deco(fn):
        global ok
        if name not in stages:
            return fn
        t = time.time()
        try:
            fn()
            print(f"[{name}] OK  {time.time()-t:.2f}s", flush=True)
        except Exception:
            ok = False
            print(f"[{name}] FAIL {time.time()-t:.2f}s", flush=True)
            traceback.print_exc()
        return fn
    return deco


@stage("imad")
def _():
    a, b = ctx.imad_peak()
    print(f"  imad {a/1e12:.2f} T/s   imad.wide {b/1e12:.2f} T/s")


@stage("field")
def _():
    for F, mod, pack in (("fq", P, cport.fqs), ("fr", R, cport.frs)):
        xs = [random.randrange(mod) for _ in range(4000)] + [0, 1, mod - 1, mod - 1, 2**253, mod - 2]
        ys = [random.randrange(mod) for _ in range(4000)] + [mod - 1, mod - 1, mod - 1, 1, 2**253, mod - 2]
        assert cport.unfr(ctx.field_op(F, "mul", pack(xs), pack(ys))) == [a * b % mod for a, b in zip(xs, ys)], F + " mul"
        assert cport.unfr(ctx.field_op(F, "add", pack(xs), pack(ys))) == [(a + b) % mod for a, b in zip(xs, ys)], F + " add"
        assert cport.unfr(ctx.field_op(F, "sub", pack(xs), pack(ys))) == [(a - b) % mod for a, b in zip(xs, ys)], F + " sub"
    try:
        ctx.field_op("fr", "mul", (R).to_bytes(32, "little"), bytes(32))
        raise AssertionError("non-canonical accepted")
    except ob.OwshenB200Error as e:
        assert e.code == -2


def _root64(leaves):
    cur = cport.unfr(leaves)
    while len(cur) > 1:
        cur = [mimc7.hash2(cur[2 * i], cur[2 * i + 1]) for i in range(len(cur) // 2)]
    return bn.fr_to_bytes(cur[0])


@stage("mimc")
def _():
    xs = [random.randrange(R) for _ in range(64)]; ys = [random.randrange(R) for _ in range(64)]
    out = cport.unfr(ctx.mimc7_hash2(cport.frs(xs), cport.frs(ys)))
    assert out == [mimc7.hash2(a, b) for a, b in zip(xs, ys)]
    n, depth = 4096, 32
    leaves = os.urandom(31 * n); leaves = b"".join(leaves[31*i:31*i+31] + b"\0" for i in range(n))
    sib = os.urandom(31 * n * depth); sib = b"".join(sib[31*i:31*i+31] + b"\0" for i in range(n * depth))
    bits = [random.randrange(1 << 32) for _ in range(n)]
    t = time.time(); got = ctx.merkle_paths(leaves, sib, bits, depth); t1 = time.time() - t
    t = time.time(); got = ctx.merkle_paths(leaves, sib, bits, depth); t2 = time.time() - t
    t = time.time(); exp = cport.merkle_paths(leaves, sib, bits, depth); t3 = time.time() - t
    assert got == exp
    print(f"  merkle 4096x32: gpu first {t1*1e3:.1f} ms, second {t2*1e3:.1f} ms (host wall), cpu oracle {t3*1e3:.0f} ms ({cport.lib().oc_num_threads()} thr)")
    lv = ctx.merkle_build(leaves[:32 * 64])
    assert lv[-32:] == _root64(leaves[:32 * 64])


@stage("ntt")
def _():
    for log_n in (0, 1, 3, 8, 10, 11, 13, 15, 18):
        n = 1 << log_n
        batch = 3 if log_n <= 15 else 1
        data = cport.frs([random.randrange(R) for _ in range(n * batch)])
        for inv in (False, True):
            for co in (False, True):
                got = ctx.ntt(data, log_n, batch, inv, co)
                exp = b"".join(cport.ntt(data[32 * n * b:32 * n * (b + 1)], inv, co) for b in range(batch))
                assert got == exp, (log_n, inv, co)
    print("  ntt sizes 2^0..2^18 ok")


def rand_points_g1(n):
    ks = cport.frs([random.randrange(R) for _ in range(n)])
    return cport.g1_fixed_mul_batch(bn.g1_to_bytes(bn.G1_GEN), ks)


def rand_points_g2(n):
    ks = cport.frs([random.randrange(R) for _ in range(n)])
    return cport.g2_fixed_mul_batch(bn.g2_to_bytes(bn.G2_GEN), ks)


@stage("msm")
def _():
    for n in (0, 1, 2, 3, 33, 255, 1024, 5000):
        pts = rand_points_g1(n)
        sc = [random.randrange(R) for _ in range(n)]
        if n >= 3: sc[0] = 0; sc[1] = 1; sc[2] = R - 1
        if n >= 33:
            pts = pts[:64 * 5] + pts[64 * 4:64 * 5] + pts[64 * 6:]          # duplicate point
            pts = pts[:64 * 7] + bytes(64) + pts[64 * 8:]                     # infinity
        got = ctx.msm_g1(pts, cport.frs(sc)); exp = cport.g1_msm(pts, cport.frs(sc))
        assert got == exp, ("g1", n)
    for n in (0, 1, 2, 77, 1024):
        pts = rand_points_g2(n)
        sc = [random.randrange(R) for _ in range(n)]
        got = ctx.msm_g2(pts, cport.frs(sc)); exp = cport.g2_msm(pts, cport.frs(sc))
        assert got == exp, ("g2", n)
    # all-equal points / equal scalars (P+P inside buckets), witness-like scalars (heavy buckets)
    n = 20000
    base = rand_points_g1(1)
    pts = base * n
    sc = cport.frs([7] * n)
    assert ctx.msm_g1(pts, sc) == cport.g1_msm(pts, sc)
    pts = rand_points_g1(n)
    sc = cport.frs([random.choice([0, 1, 1, random.randrange(1 << 64), random.randrange(R)]) for _ in range(n)])
    assert ctx.msm_g1(pts, sc) == cport.g1_msm(pts, sc)
    for log_n in (16, 20):
        n = 1 << log_n
        t = time.time(); pts = rand_points_g1(n); sc = os.urandom(31 * n); sc = b"".join(sc[31*i:31*i+31] + b"\0" for i in range(n))
        tg = time.time() - t
        t = time.time(); got = ctx.msm_g1(pts, sc); t1 = time.time() - t
        t = time.time(); got = ctx.msm_g1(pts, sc); t2 = time.time() - t
        t = time.time(); exp = cport.g1_msm(pts, sc); t3 = time.time() - t
        assert got == exp, ("g1 big", log_n)
        print(f"  msm g1 2^{log_n}: gen {tg:.1f}s gpu first {t1*1e3:.0f} ms second {t2*1e3:.0f} ms, cpu oracle {t3*1e3:.0f} ms")
    sums = rand_points_g1(9)
    exp = bytes(64)
    for i in range(9): exp = cport.g1_add(exp, sums[64 * i:64 * i + 64])
    assert ctx.g1_sum(sums) == exp
    s2 = rand_points_g2(5)
    exp = bytes(128)
    for i in range(5): exp = cport.g2_add(exp, s2[128 * i:128 * i + 128])
    assert ctx.g2_sum(s2) == exp


def rand_inputs(batch, depth):
    nul = cport.frs([random.randrange(R) for _ in range(batch)])
    sec = cport.frs([random.randrange(R) for _ in range(batch)])
    rec = cport.frs([random.randrange(1 << 160) for _ in range(batch)])
    sib = cport.frs([random.randrange(R) for _ in range(batch * depth)])
    bits = [random.randrange(1 << depth) for _ in range(batch)]
    return nul, sec, rec, sib, bits


@stage("witness")
def _():
    for depth, batch in ((2, 3), (32, 5)):
        nul, sec, rec, sib, bits = rand_inputs(batch, depth)
        got = ctx.withdraw_witness(depth, nul, sec, rec, sib, bits)
        exp = cport.withdraw_witness(nul, sec, rec, sib, bits, depth)
        assert got == exp, depth


STATE = {}


def vk_blob_from(vkb):
    return b"OGVK" + struct.pack("<II", 1, 3) + vkb["alpha1"] + vkb["beta2"] + vkb["gamma2"] + vkb["delta2"] + vkb["ic"]


def pk_blob_from(cs, pkb, depth):
    blob = b"OGPK" + struct.pack("<IIIIII", 1, depth, cs.n_constraints, cs.n_vars, cs.n_pub, pkb["log_m"])
    blob += pkb["alpha1"] + pkb["beta1"] + pkb["beta2"] + pkb["delta1"] + pkb["delta2"]
    blob += pkb["a"] + pkb["b1"] + pkb["b2"] + pkb["l"] + pkb["h"]
    for m in "AB":
        ptr, idx, val = cs.csr(m)
        blob += struct.pack("<I", len(idx)) + struct.pack(f"<{len(ptr)}I", *ptr) + struct.pack(f"<{len(idx)}I", *idx) + cport.frs(val)
    return blob


@stage("setup")
def _():
    for depth in (2, 32):
        tw = [random.randrange(1, R) for _ in range(5)]
        t = time.time(); pk, vk = ob.setup_withdraw(ctx, depth, *tw); t1 = time.time() - t
        cs = wc.build_r1cs(depth)
        t = time.time(); pkb, vkb = cport.setup_bytes(cs, *tw); t2 = time.time() - t
        assert vk == vk_blob_from(vkb), "vk"
        assert pk == pk_blob_from(cs, pkb, depth), "pk"
        print(f"  setup depth {depth}: gpu {t1:.2f}s cpu oracle {t2:.2f}s, pk {len(pk)/1e6:.1f} MB")
        STATE[depth] = (pk, vk, cs, pkb)


@stage("prove")
def _():
    for depth, batch in ((2, 3), (32, 4), (32, 70)):
        if depth not in STATE:
            tw = [random.randrange(1, R) for _ in range(5)]
            pk, vk = ob.setup_withdraw(ctx, depth, *tw)
            cs = wc.build_r1cs(depth)
            pkb, vkb = cport.setup_bytes(cs, *tw)
            STATE[depth] = (pk, vk, cs, pkb)
        pk, vk, cs, pkb = STATE[depth]
        t = time.time(); PK = ob.ProvingKey(ctx, pk); tl = time.time() - t
        nul, sec, rec, sib, bits = rand_inputs(batch, depth)
        rs = cport.frs([random.randrange(R) for _ in range(2 * batch)])
        wit = cport.withdraw_witness(nul, sec, rec, sib, bits, depth)
        opr = cport.Prover(cs, pkb)
        if batch <= 4:
            assert PK.h_evals(wit[:32 * cs.n_vars]) == opr.h_evals(wit[:32 * cs.n_vars]), "h evals"
        t = time.time(); proofs, pub = ob.prove(PK, nul, sec, rec, sib, bits, rs); t1 = time.time() - t
        t = time.time(); proofs2, _ = ob.prove(PK, nul, sec, rec, sib, bits, rs); t2 = time.time() - t
        nck = min(batch, 8)
        t = time.time(); exp = opr.prove_batch(wit[:32 * cs.n_vars * nck], rs[:64 * nck]); t3 = time.time() - t
        assert proofs == proofs2
        assert proofs[:256 * nck] == exp, "proof bytes differ from oracle"
        for i in range(min(batch, 3)):
            assert ob.verify(vk, pub[96 * i:96 * i + 96], proofs[256 * i:256 * i + 256])
        assert PK.prove_witnesses(wit[:32 * cs.n_vars * 2], rs[:128]) == proofs[:512]
        print(f"  prove depth {depth} batch {batch}: load_pk {tl:.2f}s gpu first {t1:.3f}s second {t2:.3f}s ({batch/t2:.1f} proofs/s); cpu oracle {nck} proofs {t3:.2f}s")
        PK.close()


print("launches:", ctx.launch_count)
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
