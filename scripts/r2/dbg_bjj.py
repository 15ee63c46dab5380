import random, sys, os
sys.path.insert(0, os.getcwd())
import owshen_b200 as ob
from oracle import bn254 as bn, babyjubjub as bjj
R = bn.R
ctx = ob.Context(0)
rng = random.Random(990)
n = 40
sks = [rng.randrange(R) for _ in range(n - 3)] + [12345, 0, 1]
rnds = [rng.randrange(R) for _ in range(n - 3)] + [2345, 7, 0]
msgs = [rng.randrange(R) for _ in range(n - 3)] + [123456, 9, 11]
fb = bn.fr_to_bytes
pkx, odd, sigs, st = ctx.bjj_sign_batch(b"".join(map(fb, sks)), b"".join(map(fb, rnds)), b"".join(map(fb, msgs)), 0)
exp = [bjj.to_pub(k) for k in sks]
for i in range(n):
    got = int.from_bytes(pkx[32*i:32*i+32], "little")
    e = exp[i]
    print(i, "ok" if got == (e[0] if isinstance(e, tuple) else e.x if hasattr(e,'x') else e) else ("DIFF got=%x exp=%s" % (got, e)), odd[i], st[i])
esig = [bjj.sign(k, r_, m, 0) for k, r_, m in zip(sks, rnds, msgs)]
for i in (36,37,38,39):
    print(i, sigs[96*i:96*i+96].hex()[:40], esig[i])
