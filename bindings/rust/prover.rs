// bindings/rust/prover.rs -- safe wrapper over ffi.rs in the reference's idiom (anyhow::Result as in
// /root/reference/src/blockchain/mod.rs:11; field elements are the reference's own `Fp` with little-endian
// `to_repr()` bytes, /root/reference/src/blockchain/tx/owshen_airdrop/babyjubjub/mod.rs:7-11).
// UNCOMPILED: this image has no Rust toolchain.  The executable model of the same API is owshen_b200/api.py
// (ProvingKey.prove_withdraw / verify), which the GPU parity tests exercise.
use anyhow::{anyhow, Result};
use ff::PrimeField;
use std::ffi::CStr;
use std::ptr;

use super::ffi;
use crate::blockchain::tx::owshen_airdrop::babyjubjub::Fp;

pub const DEPTH: usize = 32;

pub struct WithdrawInput {
    pub nullifier: Fp,
    pub secret: Fp,
    pub recipient: Fp,
    pub siblings: [Fp; DEPTH],
    pub path_bits: u32, // bit i = 1: the node at level i is a right child
}

#[derive(Clone)]
pub struct Proof(pub [u8; 256]); // A (G1) || B (G2) || C (G1), little-endian coordinates

pub struct PublicInputs {
    pub root: Fp,
    pub nullifier_hash: Fp,
    pub recipient: Fp,
}

/// One CUDA context + one resident proving key = one GPU.  `og_ctx` owns a single stream, so calls on one
/// `Prover` are serialised by `&mut self`; share it behind the node's `Arc<tokio::sync::Mutex<..>>`
/// (/root/reference/src/cli/node.rs:71) and call from `spawn_blocking`.
pub struct Prover {
    ctx: *mut ffi::OgCtx,
    pk: *mut ffi::OgPk,
}
unsafe impl Send for Prover {}

fn check(code: i32, ctx: *const ffi::OgCtx) -> Result<()> {
    if code == ffi::OG_OK {
        return Ok(());
    }
    let what = unsafe { CStr::from_ptr(ffi::og_strerror(code)) }.to_string_lossy().into_owned();
    let detail = if ctx.is_null() {
        String::new()
    } else {
        unsafe { CStr::from_ptr(ffi::og_last_error(ctx)) }.to_string_lossy().into_owned()
    };
    Err(anyhow!("owshen_b200: {} ({}) {}", what, code, detail))
}

fn push(buf: &mut Vec<u8>, x: &Fp) {
    buf.extend_from_slice(x.to_repr().as_ref()); // 32 bytes, little-endian, canonical
}

fn fp_from(bytes: &[u8]) -> Result<Fp> {
    let mut repr = <Fp as PrimeField>::Repr::default();
    repr.as_mut().copy_from_slice(bytes);
    Option::<Fp>::from(Fp::from_repr(repr)).ok_or_else(|| anyhow!("owshen_b200: non-canonical field element"))
}

impl Prover {
    /// `device`: CUDA ordinal; `pk_bytes`: the OGPK blob written by the setup.
    pub fn new(device: i32, pk_bytes: &[u8]) -> Result<Self> {
        let mut ctx = ptr::null_mut();
        check(unsafe { ffi::og_init(device, &mut ctx) }, ptr::null())?; // OG_E_NO_DEVICE: there is no CPU path
        let mut pk = ptr::null_mut();
        let rc = unsafe { ffi::og_load_pk(ctx, pk_bytes.as_ptr(), pk_bytes.len() as u64, &mut pk) };
        if let Err(e) = check(rc, ctx) {
            unsafe { ffi::og_free(ctx) };
            return Err(e);
        }
        Ok(Prover { ctx, pk })
    }

    /// prove(): one Groth16 proof per input; `rs[i]` = the (r, s) blinding pair of proof i (draw from OsRng).
    pub fn prove(&mut self, inputs: &[WithdrawInput], rs: &[(Fp, Fp)]) -> Result<Vec<(Proof, PublicInputs)>> {
        if inputs.len() != rs.len() {
            return Err(anyhow!("owshen_b200: one (r, s) pair per proof"));
        }
        let n = inputs.len();
        let (mut nul, mut sec, mut rcp) = (Vec::with_capacity(32 * n), Vec::with_capacity(32 * n), Vec::with_capacity(32 * n));
        let (mut sib, mut blind) = (Vec::with_capacity(32 * DEPTH * n), Vec::with_capacity(64 * n));
        let mut bits = Vec::with_capacity(n);
        for (w, (r, s)) in inputs.iter().zip(rs) {
            push(&mut nul, &w.nullifier);
            push(&mut sec, &w.secret);
            push(&mut rcp, &w.recipient);
            for x in &w.siblings {
                push(&mut sib, x);
            }
            bits.push(w.path_bits);
            push(&mut blind, r);
            push(&mut blind, s);
        }
        let mut proofs = vec![0u8; 256 * n];
        let mut public = vec![0u8; 96 * n];
        let rc = unsafe {
            ffi::og_groth16_prove_withdraw(
                self.ctx, self.pk, nul.as_ptr(), sec.as_ptr(), rcp.as_ptr(), sib.as_ptr(), bits.as_ptr(), n as u32,
                blind.as_ptr(), proofs.as_mut_ptr(), public.as_mut_ptr(),
            )
        };
        check(rc, self.ctx)?;
        let mut out = Vec::with_capacity(n);
        for i in 0..n {
            let mut p = [0u8; 256];
            p.copy_from_slice(&proofs[256 * i..256 * (i + 1)]);
            let q = &public[96 * i..96 * (i + 1)];
            out.push((
                Proof(p),
                PublicInputs { root: fp_from(&q[0..32])?, nullifier_hash: fp_from(&q[32..64])?, recipient: fp_from(&q[64..96])? },
            ));
        }
        Ok(out)
    }

    /// MiMC7 two-to-one hashes, one launch for the whole slice: the node of a Merkle level from its children.
    pub fn hash2(&mut self, left: &[Fp], right: &[Fp]) -> Result<Vec<Fp>> {
        if left.len() != right.len() {
            return Err(anyhow!("owshen_b200: left/right length mismatch"));
        }
        let (mut l, mut r) = (Vec::with_capacity(32 * left.len()), Vec::with_capacity(32 * left.len()));
        left.iter().for_each(|x| push(&mut l, x));
        right.iter().for_each(|x| push(&mut r, x));
        let mut out = vec![0u8; 32 * left.len()];
        check(unsafe { ffi::og_mimc7_hash2(self.ctx, l.as_ptr(), r.as_ptr(), left.len() as u64, out.as_mut_ptr()) }, self.ctx)?;
        out.chunks(32).map(fp_from).collect()
    }
}

impl Drop for Prover {
    fn drop(&mut self) {
        unsafe {
            ffi::og_free_pk(self.pk); // the key's tables are device memory of this context's GPU: free it first
            ffi::og_free(self.ctx);
        }
    }
}

/// verify(): host-side pairing check; Ok(false) for a well-formed proof that does not verify.
pub fn verify(vk: &[u8], public: &PublicInputs, proof: &Proof) -> Result<bool> {
    let mut p = Vec::with_capacity(96);
    push(&mut p, &public.root);
    push(&mut p, &public.nullifier_hash);
    push(&mut p, &public.recipient);
    match unsafe { ffi::og_groth16_verify(vk.as_ptr(), vk.len() as u64, p.as_ptr(), 3, proof.0.as_ptr()) } {
        ffi::OG_OK => Ok(true),
        ffi::OG_E_VERIFY => Ok(false),
        e => check(e, ptr::null()).map(|_| false),
    }
}
