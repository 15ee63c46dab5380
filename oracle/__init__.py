"""CPU oracle for the owshen-b200 Groth16 hot path.  TEST INFRASTRUCTURE ONLY.

This package is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product (``owshen_b200``) never
imports anything from here and fails loudly when its CUDA library is missing.

PARITY UNPINNED: the reference snapshot (OwshenNetwork/owshen @ c7b1f00) holds no
Groth16 prover, MSM, NTT, MiMC or Merkle tree, so there
is no reference output to pin against.  The only conventions the reference
fixes are the BN254 scalar field, its little-endian 32-byte representation and
its multiplicative generator 7
(src/blockchain/tx/owshen_airdrop/babyjubjub/mod.rs:7-11) plus the BabyJubJub
constants (same file :174-189); both are honoured and tested here.  Everything
else restates public standards (alt_bn128 / EIP-196/197, Groth16 [Groth,
EUROCRYPT 2016], circomlib MiMC7) and is pinned by (i) two independently
written implementations that must agree bit-for-bit -- the pure-Python
big-integer spec in this package and the C port in ``oracle/cpu`` -- and
(ii) algebraic invariants (pairing equation, NTT vs naive DFT, MSM linearity).
"""
