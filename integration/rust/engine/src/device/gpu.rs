//! Replacement for engine/src/device/gpu.rs (cudarc + NVRTC + cuBLAS) — a thin shim over the
//! rama_b200 C ABI.  `GPU` keeps its name, `new()`, `Send + Sync` and the full `Device<T>` impl, so
//! engine/src/main.rs:70-98 and engine/src/lib.rs:99-152 compile unchanged.
//!
//! Storage: `DevBuf` (below) takes the place of `cudarc::driver::CudaSlice<f32>`.  A RunState
//! built for the GPU additionally owns a `rama_session` (KV cache, step graph); `forward()` reaches
//! it through the new defaulted trait method `forward_fused` and otherwise
//! falls back to the per-op body of infer.rs, which this impl also serves.
use std::sync::Mutex;

use super::device::Device;
use super::ffi::*;
use crate::transformer::state::{RunState, RunStateView};
use crate::transformer::{Config, MutView, Storage, View};

pub struct DevBuf {
    pub ptr: *mut f32,
    pub len: usize,
    ctx: *mut rama_ctx,
    /// Some(session) only on RunState.x: the fused path's handle (one session per RunState).
    pub session: Option<*mut rama_session>,
}
unsafe impl Send for DevBuf {}
unsafe impl Sync for DevBuf {}
impl Storage for DevBuf {
    fn length(&self) -> usize { self.len }
}
impl Drop for DevBuf {
    fn drop(&mut self) {
        unsafe {
            if let Some(s) = self.session.take() { rama_session_destroy(s); }
            rama_dev_free(self.ctx, self.ptr);
        }
    }
}

#[derive(Debug)]
pub struct GPU {
    pub ctx: *mut rama_ctx,
    lock: Mutex<()>, // op-level calls share the context's op stream
}
unsafe impl Send for GPU {}
unsafe impl Sync for GPU {}

impl GPU {
    /// `RAMA_GPUS=N` makes this one handle tensor-parallel over devices 0..N-1 of the process
    /// (rama_ctx_create_multi): main.rs and lib.rs keep calling `GPU::new()` unchanged.
    pub fn new() -> Self {
        let n: i32 = std::env::var("RAMA_GPUS").ok().and_then(|v| v.parse().ok()).unwrap_or(1);
        let mut ctx = std::ptr::null_mut();
        if n > 1 {
            ck(unsafe { rama_ctx_create_multi(n, std::ptr::null(), &mut ctx) });
        } else {
            ck(unsafe { rama_ctx_create(0, std::ptr::null(), &mut ctx) });
        }
        Self { ctx, lock: Mutex::new(()) }
    }
    pub fn alloc(&self, host: &[f32]) -> DevBuf {
        let mut p = std::ptr::null_mut();
        ck(unsafe { rama_dev_alloc(self.ctx, host.len(), &mut p) });
        ck(unsafe { rama_dev_h2d(self.ctx, p, host.as_ptr(), host.len()) });
        DevBuf { ptr: p, len: host.len(), ctx: self.ctx, session: None }
    }
    pub fn cfg(c: &Config) -> rama_config {
        rama_config { dim: c.dim as i32, hidden_dim: c.hidden_dim as i32, n_layers: c.n_layers as i32,
            n_heads: c.n_heads as i32, n_kv_heads: c.n_kv_heads as i32, vocab_size: c.vocab_size as i32,
            seq_len: c.seq_len as i32, shared_weight: c.shared_weight as i32 }
    }
}
impl Drop for GPU {
    fn drop(&mut self) { unsafe { rama_ctx_destroy(self.ctx); } }
}

#[inline] fn p(v: &View<'_, DevBuf>) -> *const f32 { unsafe { v.data.ptr.add(v.range.start) as *const f32 } }
#[inline] fn pm(v: &MutView<'_, DevBuf>) -> *mut f32 { unsafe { v.data.ptr.add(v.range.start) } }

impl Device<DevBuf> for GPU {
    fn array_add(&self, t: &mut MutView<'_, DevBuf>, s: &View<'_, DevBuf>, n: usize) {
        let _g = self.lock.lock().unwrap();
        ck(unsafe { rama_op_array_add(self.ctx, pm(t), p(s), n) });
    }
    fn array_mult(&self, t: &mut MutView<'_, DevBuf>, s: &View<'_, DevBuf>, n: usize) {
        let _g = self.lock.lock().unwrap();
        ck(unsafe { rama_op_array_mult(self.ctx, pm(t), p(s), n) });
    }
    fn sinu(&self, o: &mut MutView<'_, DevBuf>, n: usize) {
        let _g = self.lock.lock().unwrap();
        ck(unsafe { rama_op_sinu(self.ctx, pm(o), n) });
    }
    fn multi_head_attention(&self, rsv: &mut RunStateView<'_, DevBuf>, cfg: &Config, layer: usize, pos: usize) {
        let _g = self.lock.lock().unwrap();
        let c = GPU::cfg(cfg);
        ck(unsafe { rama_op_multi_head_attention(self.ctx, pm(&rsv.xb), pm(&rsv.att), pm(&rsv.q) as *const f32,
            pm(&rsv.key_cache) as *const f32, pm(&rsv.value_cache) as *const f32, &c, layer as i32, pos as i32) });
    }
    fn copy_from_slice(&self, t: &mut MutView<'_, DevBuf>, s: &View<'_, DevBuf>, n: usize) {
        let _g = self.lock.lock().unwrap();
        ck(unsafe { rama_op_copy_from_slice(self.ctx, pm(t), p(s), n) });
    }
    fn rmsnorm(&self, o: &mut MutView<'_, DevBuf>, x: &View<'_, DevBuf>, w: &View<'_, DevBuf>, n: usize) {
        let _g = self.lock.lock().unwrap();
        ck(unsafe { rama_op_rmsnorm(self.ctx, pm(o), p(x), p(w), n) });
    }
    fn apply_position(&self, q: &mut MutView<'_, DevBuf>, k: &mut MutView<'_, DevBuf>, pr: &View<'_, DevBuf>,
                      pi: &View<'_, DevBuf>, head_size: usize) {
        let _g = self.lock.lock().unwrap();
        ck(unsafe { rama_op_apply_position(self.ctx, pm(q), pm(k), p(pr), p(pi), head_size) });
    }
    fn matmul(&self, o: &mut MutView<'_, DevBuf>, a: &View<'_, DevBuf>, b: &View<'_, DevBuf>, width: usize,
              o_rows: usize, o_cols: usize) {
        let _g = self.lock.lock().unwrap();
        ck(unsafe { rama_op_matmul(self.ctx, pm(o), p(a), p(b), width, o_rows, o_cols) });
    }
    fn softmax<'a>(&self, x: &mut MutView<'a, DevBuf>, n: usize) {
        let _g = self.lock.lock().unwrap();
        ck(unsafe { rama_op_softmax(self.ctx, pm(x), n) });
    }
    fn sample<'a>(&self, cfg: &Config, rsv: &mut RunStateView<'a, DevBuf>, temperature: f32, topp: f32) -> usize {
        let mut next = 0i32;
        match rsv.x.data.session {
            // fused path: logits live in the session; 8 bytes cross PCIe
            Some(s) => ck(unsafe { rama_sample(s, temperature, topp, &mut next) }),
            None => {
                let _g = self.lock.lock().unwrap();
                ck(unsafe { rama_op_sample(self.ctx, pm(&rsv.logits), cfg.vocab_size, temperature, topp, &mut next) })
            }
        }
        next as usize
    }
    fn to_cpu(&self, state: &RunStateView<DevBuf>, cpu: &mut RunState<Vec<f32>>) {
        let bufs: [(&MutView<DevBuf>, &mut Vec<f32>); 12] = [
            (&state.x, &mut cpu.x), (&state.xb, &mut cpu.xb), (&state.xb2, &mut cpu.xb2), (&state.hb, &mut cpu.hb),
            (&state.hb2, &mut cpu.hb2), (&state.q, &mut cpu.q), (&state.k, &mut cpu.k), (&state.v, &mut cpu.v),
            (&state.att, &mut cpu.att), (&state.logits, &mut cpu.logits), (&state.key_cache, &mut cpu.key_cache),
            (&state.value_cache, &mut cpu.value_cache)];
        for (i, (d, h)) in bufs.into_iter().enumerate() {
            match state.x.data.session {
                Some(s) => { let mut n = 0usize; ck(unsafe { rama_state_to_host(s, i as i32, h.as_mut_ptr(), h.len(), &mut n) }); }
                None => ck(unsafe { rama_dev_d2h(self.ctx, h.as_mut_ptr(), pm(d) as *const f32, h.len()) }),
            }
        }
    }
    // New defaulted trait method (device.rs): the fused, graph-replayed step.
    fn forward_fused<'a>(&self, rsv: &mut RunStateView<'a, DevBuf>, token: usize, pos: usize) -> bool {
        match rsv.x.data.session {
            Some(s) => { ck(unsafe { rama_forward(s, token as i32, pos as i32) }); true }
            None => false,
        }
    }
}
