"""owshen_b200 -- B200-native (sm_100a) Groth16 backend for privacy-pool withdraw proofs over BN254.

Python is the host language here because the reference's (Rust) toolchain is absent from this image;
everything below is a thin ctypes veneer over the C ABI in include/owshen_b200.h, which is the real
drop-in boundary (INTEGRATION.md shows the Rust binding).  There is no CPU fallback: importing works
anywhere, but creating a Context without a CUDA device raises.
"""
from .kvstore import KvStore, MirrorKvStore, RamKvStore
from .api import (Context, ProvingKey, MerkleTree, OwshenB200Error, lib, build_library, prove, verify,
                  setup_withdraw, FR_MODULUS, PROOF_BYTES)

__all__ = ["Context", "ProvingKey", "MerkleTree", "OwshenB200Error", "lib", "build_library", "prove", "verify",
           "setup_withdraw", "FR_MODULUS", "PROOF_BYTES", "KvStore", "RamKvStore", "MirrorKvStore"]
