"""MiMC7 (circomlib flavour) over BN254 Fr, multi-hash and Merkle paths.

EXTERNAL standard, not from the reference (its only field "hash" is a
placeholder product, babyjubjub/mod.rs:202-204).  Restated from the published
circomlib algorithm (src/mimc7.js): 91 rounds of x -> (x + k + c_i)^7 with
c_0 = 0 and c_i = keccak256^(i+1)("mimc") mod r; hash(x, k) = perm(x, k) + k;
multi_hash(xs, key): r = key; for x in xs: r = r + x + hash(x, r).
Pinned by the two circomlib constants every copy of mimc7.circom lists first
(c_1 = 2088896141...570981, c_2 = 1526512611...317193), checked in tests.
"""
from .keccak import keccak256
from .bn254 import R

N_ROUNDS = 91
SEED = b"mimc"


def round_constants(n_rounds: int = N_ROUNDS):
    cts = [0] * n_rounds
    c = keccak256(SEED)
    for i in range(1, n_rounds):
        c = keccak256(c)
        cts[i] = int.from_bytes(c, "big") % R
    return cts


CONSTANTS = round_constants()


def mimc7_hash(x: int, k: int, n_rounds: int = N_ROUNDS) -> int:
    r = x % R
    for i in range(n_rounds):
        t = (r + k + CONSTANTS[i]) % R        # c_0 = 0
        r = pow(t, 7, R)
    return (r + k) % R


def multi_hash(xs, key: int = 0, n_rounds: int = N_ROUNDS) -> int:
    r = key % R
    for x in xs:
        r = (r + x + mimc7_hash(x, r, n_rounds)) % R
    return r


def hash2(left: int, right: int, n_rounds: int = N_ROUNDS) -> int:
    """The Merkle node function: MultiMiMC7(2 inputs, key 0)."""
    return multi_hash([left, right], 0, n_rounds)


def merkle_path_nodes(leaf: int, siblings, path_bits, n_rounds: int = N_ROUNDS):
    """All depth+1 nodes from leaf to root.  bit=1 means the current node is the
    RIGHT child at that level (sibling on the left)."""
    nodes = [leaf % R]
    cur = leaf % R
    for sib, bit in zip(siblings, path_bits):
        left, right = (sib, cur) if bit else (cur, sib)
        cur = hash2(left, right, n_rounds)
        nodes.append(cur)
    return nodes


class MerkleTree:
    """Sparse fixed-depth MiMC7 Merkle tree with zero-subtree defaults."""

    def __init__(self, depth: int, n_rounds: int = N_ROUNDS):
        self.depth, self.n_rounds = depth, n_rounds
        self.zeros = [0]
        for _ in range(depth):
            self.zeros.append(hash2(self.zeros[-1], self.zeros[-1], n_rounds))
        self.levels = [dict() for _ in range(depth + 1)]
        self.n_leaves = 0

    def _get(self, lvl, idx):
        return self.levels[lvl].get(idx, self.zeros[lvl])

    def insert(self, leaf: int) -> int:
        idx = self.n_leaves
        self.n_leaves += 1
        self.levels[0][idx] = leaf % R
        i = idx
        for lvl in range(self.depth):
            l, r = self._get(lvl, i & ~1), self._get(lvl, i | 1)
            i >>= 1
            self.levels[lvl + 1][i] = hash2(l, r, self.n_rounds)
        return idx

    def root(self) -> int:
        return self._get(self.depth, 0)

    def path(self, idx: int):
        sibs, bits = [], []
        i = idx
        for lvl in range(self.depth):
            sibs.append(self._get(lvl, i ^ 1))
            bits.append(i & 1)
            i >>= 1
        return sibs, bits
