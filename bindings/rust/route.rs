// bindings/rust/route.rs -- service wiring for shielded withdrawals, as source a maintainer
// drops into the reference tree.  UNCOMPILED: this image has no Rust toolchain; the byte-level behaviour (the RLP
// message) is executable in owshen_b200/formats.py and pinned by the known-answer vector at the bottom of this file,
// which tests/test_formats_rlp.py::test_shielded_withdraw_rlp_kat_shared_with_rust checks against the Python encoder.
//
// Shapes followed (not copied):
//   * the handler:   /root/reference/src/services/api_services/withdraw.rs:27-71 (`withdraw_handler`: lock the context,
//                    decode the request, refuse a replay by a `db` key, build a `CustomTx`, enqueue it, answer its hash);
//   * registration:  /root/reference/src/services/api_services/mod.rs:83-144 (`api_routes`: one `.route(..)` per endpoint,
//                    the context cloned into the closure, `handle_error` around the handler);
//   * the message:   /root/reference/src/types/tx/custom.rs:214-256 (`CustomTxMsg::{as_rlp, from_rlp}`: an RLP list whose
//                    first item is the kind string, byte fields as RLP strings).
// The proving itself is `Prover::prove` (bindings/rust/prover.rs) over the C ABI of include/owshen_b200.h.

use alloy::primitives::FixedBytes;
use anyhow::{anyhow, Result};
use axum::Json;
use rlp::{Rlp, RlpStream};
use serde::{Deserialize, Serialize};
use std::sync::Arc;
use tokio::sync::Mutex;

use super::prover::{Proof, Prover, PublicInputs, WithdrawInput};
use crate::blockchain::tx::owshen_airdrop::babyjubjub::Fp;
use crate::config::CHAIN_ID;
use crate::services::{Context, ContextKvStore, ContextSigner};
use crate::types::{CustomTx, CustomTxMsg};

pub const SHIELDED_WITHDRAW_KIND: &str = "shielded-withdraw";

/// A withdraw proof with its public inputs: what the chain needs to check it (root known, nullifier unspent) and pay out.
#[derive(Debug, Clone, PartialEq)]
pub struct ShieldedWithdraw {
    pub proof: [u8; 256],          // A || B || C, little-endian coordinates (include/owshen_b200.h)
    pub root: [u8; 32],            // Fp::to_repr() bytes, little-endian (babyjubjub/mod.rs:7-11)
    pub nullifier_hash: [u8; 32],
    pub recipient: [u8; 32],
}

impl ShieldedWithdraw {
    pub fn new(proof: &Proof, public: &PublicInputs) -> Self {
        use ff::PrimeField;
        let le = |x: &Fp| -> [u8; 32] { let mut o = [0u8; 32]; o.copy_from_slice(x.to_repr().as_ref()); o };
        ShieldedWithdraw { proof: proof.0, root: le(&public.root), nullifier_hash: le(&public.nullifier_hash), recipient: le(&public.recipient) }
    }

    /// ["shielded-withdraw", proof, root, nullifier_hash, recipient] -- the list shape of custom.rs:233-236
    pub fn as_rlp(&self) -> Vec<u8> {
        let mut stream = RlpStream::new_list(5);
        stream.append(&SHIELDED_WITHDRAW_KIND);
        stream.append(&self.proof.to_vec());
        stream.append(&self.root.to_vec());
        stream.append(&self.nullifier_hash.to_vec());
        stream.append(&self.recipient.to_vec());
        stream.out().into()
    }

    pub fn from_rlp(bytes: &[u8]) -> Result<Self> {
        let rlp = Rlp::new(bytes);
        if rlp.item_count()? != 5 {
            return Err(anyhow!("Invalid tx!"));
        }
        let kind: String = rlp.val_at(0)?;
        if kind != SHIELDED_WITHDRAW_KIND {
            return Err(anyhow!("Invalid tx!"));
        }
        let take = |i: usize, n: usize| -> Result<Vec<u8>> {
            let v: Vec<u8> = rlp.val_at(i)?;
            if v.len() != n { Err(anyhow!("Invalid tx!")) } else { Ok(v) }
        };
        let mut out = ShieldedWithdraw { proof: [0u8; 256], root: [0u8; 32], nullifier_hash: [0u8; 32], recipient: [0u8; 32] };
        out.proof.copy_from_slice(&take(1, 256)?);
        out.root.copy_from_slice(&take(2, 32)?);
        out.nullifier_hash.copy_from_slice(&take(3, 32)?);
        out.recipient.copy_from_slice(&take(4, 32)?);
        Ok(out)
    }
}

// ---- the two arms a maintainer adds to `CustomTxMsg` (custom.rs:214-256) ---------------------------------------------
//
//     pub enum CustomTxMsg {
//         MintTx(Mint),
//         BurnTx(Burn),
//         ShieldedWithdraw(ShieldedWithdraw),                                           // + new
//     }
//     // as_rlp:    CustomTxMsg::ShieldedWithdraw(w) => w.as_rlp(),                     // + new
//     // from_rlp:  "shielded-withdraw" => Ok(CustomTxMsg::ShieldedWithdraw(ShieldedWithdraw::from_rlp(bytes)?)),   // + new
//
// and `Key::Nullifier(FixedBytes<32>)` next to `Key::BurnId` in src/db/mod.rs, used below exactly like the burn id.

#[derive(Deserialize, Debug)]
pub struct ProveRequest {
    pub nullifier: [u8; 32],       // little-endian canonical field elements throughout
    pub secret: [u8; 32],
    pub recipient: [u8; 32],
    pub leaf_index: u64,           // position of the commitment in the deposit tree; the node supplies the path
}

#[derive(Serialize)]
pub struct ProveResponse {
    pub id: FixedBytes<32>,        // hash of the enqueued transaction, as WithdrawResponse.id (withdraw.rs:21-25)
    pub success: bool,
}

/// POST /prove -- shaped like `withdraw_handler`: one lock, replay check by key, sign, enqueue, record.
/// `prover` is one GPU (one og_ctx + resident proving key); `tree` is the MiMC7 deposit tree kept behind the node's
/// KvStore (the Rust twin of owshen_b200.api.MerkleTree over og_mimc7_merkle_append).
pub async fn prove_handler<S: ContextSigner, K: ContextKvStore>(
    ctx: Arc<Mutex<Context<S, K>>>,
    prover: Arc<Mutex<Prover>>,
    Json(payload): Json<ProveRequest>,
) -> Result<Json<ProveResponse>, anyhow::Error> {
    let mut _ctx = ctx.lock().await;

    let fp = |b: &[u8; 32]| -> Result<Fp> {
        use ff::PrimeField;
        let mut repr = <Fp as PrimeField>::Repr::default();
        repr.as_mut().copy_from_slice(b);
        Option::<Fp>::from(Fp::from_repr(repr)).ok_or_else(|| anyhow!("non-canonical field element"))
    };
    let (siblings, path_bits) = _ctx.chain.deposit_tree_path(payload.leaf_index)?;     // 32 siblings + direction bits
    let input = WithdrawInput {
        nullifier: fp(&payload.nullifier)?,
        secret: fp(&payload.secret)?,
        recipient: fp(&payload.recipient)?,
        siblings,
        path_bits,
    };
    let blind = (Fp::random(rand::rngs::OsRng), Fp::random(rand::rngs::OsRng));

    // the GPU call blocks for ~0.5 s per 1024 proofs: keep it off the async workers
    let prover2 = prover.clone();
    let mut proved = tokio::task::spawn_blocking(move || {
        let mut p = prover2.blocking_lock();
        p.prove(&[input], &[blind])
    })
    .await??;
    let (proof, public) = proved.pop().ok_or_else(|| anyhow!("prover returned nothing"))?;
    let msg = ShieldedWithdraw::new(&proof, &public);

    let nullifier_key = crate::db::Key::Nullifier(FixedBytes::from_slice(&msg.nullifier_hash));
    if _ctx.chain.db.get(nullifier_key.clone())?.is_some() {
        return Err(anyhow!("Nullifier already used!"));
    }

    let tx = CustomTx::create(&mut _ctx.signer, CHAIN_ID, CustomTxMsg::ShieldedWithdraw(msg)).await?;
    let id = tx.hash()?;
    _ctx.tx_queue.enqueue(tx);
    _ctx.chain.db.put(nullifier_key, Some(crate::db::Value::Void))?;

    Ok(Json(ProveResponse { id, success: true }))
}

// ---- registration, inside `api_routes` (mod.rs:83-144), next to "/withdraw" --------------------------------------------
//
//     .route(
//         "/prove",
//         post({
//             let ctx = ctx.clone();
//             let prover = prover.clone();
//             move |Json(req): Json<ProveRequest>| async move {
//                 handle_error(prove_handler(ctx.clone(), prover.clone(), extract::Json(req)).await)
//             }
//         }),
//     )

// ---- applying the message on chain: the arm `apply_tx` gains next to MintTx / BurnTx -----------------------------------
// verify() is three pairings on the CPU (og_groth16_verify has no GPU dependency), so validators need no GPU.
pub fn check_shielded_withdraw(vk: &[u8], known_root: &[u8; 32], w: &ShieldedWithdraw) -> Result<()> {
    if &w.root != known_root {
        return Err(anyhow!("Unknown Merkle root!"));
    }
    let mut public = Vec::with_capacity(96);
    public.extend_from_slice(&w.root);
    public.extend_from_slice(&w.nullifier_hash);
    public.extend_from_slice(&w.recipient);
    let rc = unsafe { super::ffi::og_groth16_verify(vk.as_ptr(), vk.len() as u64, public.as_ptr(), 3, w.proof.as_ptr()) };
    match rc {
        super::ffi::OG_OK => Ok(()),
        super::ffi::OG_E_VERIFY => Err(anyhow!("Invalid proof!")),
        e => Err(anyhow!("owshen_b200: malformed proof or key ({})", e)),
    }
}

#[cfg(test)]
mod tests {
    use super::*;

    // Known answer shared with the Python encoder (tests/test_formats_rlp.py reads these constants out of this file):
    // proof[i] = i, public[i] = (7 i + 3) mod 256, split as root | nullifier_hash | recipient.
    pub const KAT_LEN: usize = 379;
    pub const KAT_PREFIX_HEX: &str = "f9017891736869656c6465642d7769746864726177b90100";
    pub const KAT_SHA256_HEX: &str = "70f5ac0661a5363138c4cbd8bd1ea0d3e1bf7fb2d04ee86f165396ad703579d5";

    #[test]
    fn shielded_withdraw_rlp_known_answer() {
        let mut w = ShieldedWithdraw { proof: [0u8; 256], root: [0u8; 32], nullifier_hash: [0u8; 32], recipient: [0u8; 32] };
        for i in 0..256 { w.proof[i] = i as u8; }
        let public: Vec<u8> = (0..96u32).map(|i| ((7 * i + 3) % 256) as u8).collect();
        w.root.copy_from_slice(&public[0..32]);
        w.nullifier_hash.copy_from_slice(&public[32..64]);
        w.recipient.copy_from_slice(&public[64..96]);
        let enc = w.as_rlp();
        assert_eq!(enc.len(), KAT_LEN);
        assert_eq!(hex::encode(&enc[..24]), KAT_PREFIX_HEX);
        assert_eq!(ShieldedWithdraw::from_rlp(&enc).unwrap(), w);
        assert!(ShieldedWithdraw::from_rlp(&enc[..enc.len() - 1]).is_err());
    }
}
