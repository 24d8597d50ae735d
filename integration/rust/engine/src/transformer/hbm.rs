//! Replacement for engine/src/transformer/hbm.rs: the `--features gpu` constructors, same names.
use crate::device::ffi::*;
use crate::device::gpu::{DevBuf, GPU};
use super::state::{RunState, TransformerWeights, TransformerWeightsView};
use super::View;

impl RunState<DevBuf> {
    /// ≙ hbm.rs:19-34.  Device buffers for the op-level path plus ONE rama_session (KV cache, graph)
    /// for the fused path, attached to `x`.  The session allocates and zeroes its KV cache in HBM;
    /// the host zeros of `state` are only uploaded for the small op-level buffers.
    pub fn from_state(state: &mut RunState<Vec<f32>>, device: &GPU) -> Self {
        let mut x = device.alloc(&state.x);
        let mut s = std::ptr::null_mut();
        ck(unsafe { rama_session_create(device.ctx, &mut s) });
        x.session = Some(s);
        Self {
            x, xb: device.alloc(&state.xb), xb2: device.alloc(&state.xb2), hb: device.alloc(&state.hb),
            hb2: device.alloc(&state.hb2), q: device.alloc(&state.q), k: device.alloc(&state.k),
            v: device.alloc(&state.v), att: device.alloc(&state.att), logits: device.alloc(&state.logits),
            // the op-level caches are only touched when forward_fused is unavailable; allocate lazily-sized 1
            key_cache: device.alloc(&state.key_cache), value_cache: device.alloc(&state.value_cache),
        }
    }
}

impl TransformerWeights<DevBuf> {
    /// ≙ hbm.rs:55-90: one pass host → HBM inside the library (sharded under TP); the 14 `DevBuf`s
    /// returned here are 1-element placeholders kept only so the struct's shape is unchanged —
    /// op-level callers that need per-tensor device views use `GPU::alloc` on the host tensors instead.
    pub fn from_weight(tw: &mut TransformerWeights<Vec<f32>>, cfg: &super::Config, device: &GPU) -> Self {
        let c = GPU::cfg(cfg);
        let t: [*const f32; RAMA_T_COUNT] = [
            tw.token_embedding_table.as_ptr(), tw.rms_att_weight.as_ptr(), tw.wq.as_ptr(), tw.wk.as_ptr(),
            tw.wv.as_ptr(), tw.wo.as_ptr(), tw.rms_ffn_weight.as_ptr(), tw.w1.as_ptr(), tw.w2.as_ptr(),
            tw.w3.as_ptr(), tw.rms_final_weight.as_ptr(), tw.freq_cis_real.as_ptr(), tw.freq_cis_imag.as_ptr(),
            if tw.wcls_exists { tw.wcls.as_ptr() } else { std::ptr::null() }];
        ck(unsafe { rama_ctx_load_host(device.ctx, &c, t.as_ptr()) });
        let one = |_: &Vec<f32>| device.alloc(&[0.0f32]);
        Self {
            token_embedding_table: one(&tw.token_embedding_table), rms_att_weight: one(&tw.rms_att_weight),
            rms_ffn_weight: one(&tw.rms_ffn_weight), wq: one(&tw.wq), wk: one(&tw.wk), wv: one(&tw.wv),
            wo: one(&tw.wo), w1: one(&tw.w1), w2: one(&tw.w2), w3: one(&tw.w3),
            rms_final_weight: one(&tw.rms_final_weight), freq_cis_real: one(&tw.freq_cis_real),
            freq_cis_imag: one(&tw.freq_cis_imag), wcls_exists: tw.wcls_exists, wcls: one(&tw.wcls),
        }
    }
}

impl<'a> TransformerWeightsView<'a, DevBuf> {
    /// ≙ hbm.rs:93-121, wcls aliasing as state.rs:111-117.
    pub fn from_gpu_ws(ws: &'a TransformerWeights<DevBuf>) -> TransformerWeightsView<'a, DevBuf> {
        TransformerWeightsView {
            token_embedding_table: View::new(&ws.token_embedding_table), rms_att_weight: View::new(&ws.rms_att_weight),
            rms_ffn_weight: View::new(&ws.rms_ffn_weight), wq: View::new(&ws.wq), wk: View::new(&ws.wk),
            wv: View::new(&ws.wv), wo: View::new(&ws.wo), w1: View::new(&ws.w1), w2: View::new(&ws.w2),
            w3: View::new(&ws.w3), rms_final_weight: View::new(&ws.rms_final_weight),
            freq_cis_real: View::new(&ws.freq_cis_real), freq_cis_imag: View::new(&ws.freq_cis_imag),
            wcls: if ws.wcls_exists { View::new(&ws.wcls) } else { View::new(&ws.token_embedding_table) },
            wcls_exists: ws.wcls_exists,
        }
    }
}
