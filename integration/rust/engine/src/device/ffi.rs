//! Raw bindings of include/rama_b200.h (hand-written; the header is small and stable, so no bindgen).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_float, c_int};

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rama_config {
    pub dim: i32, pub hidden_dim: i32, pub n_layers: i32, pub n_heads: i32,
    pub n_kv_heads: i32, pub vocab_size: i32, pub seq_len: i32, pub shared_weight: i32,
}
#[repr(C)]
pub struct rama_tp { pub rank: i32, pub world: i32, pub nccl_id: [u8; 128] }
#[repr(C)] pub struct rama_ctx { _p: [u8; 0] }
#[repr(C)] pub struct rama_session { _p: [u8; 0] }

pub const RAMA_T_COUNT: usize = 14;

extern "C" {
    pub fn rama_last_error() -> *const c_char;
    pub fn rama_ctx_create(device: c_int, tp: *const rama_tp, out: *mut *mut rama_ctx) -> c_int;
    pub fn rama_ctx_create_multi(n_gpus: i32, devices: *const i32, out: *mut *mut rama_ctx) -> c_int;
    pub fn rama_ctx_destroy(ctx: *mut rama_ctx) -> c_int;
    pub fn rama_ctx_load_file(ctx: *mut rama_ctx, path: *const c_char) -> c_int;
    pub fn rama_ctx_load_host(ctx: *mut rama_ctx, cfg: *const rama_config, tensors: *const *const c_float) -> c_int;
    pub fn rama_session_create(ctx: *mut rama_ctx, out: *mut *mut rama_session) -> c_int;
    pub fn rama_session_destroy(s: *mut rama_session) -> c_int;
    pub fn rama_forward(s: *mut rama_session, token: i32, pos: i32) -> c_int;
    pub fn rama_sample(s: *mut rama_session, temperature: c_float, topp: c_float, next: *mut i32) -> c_int;
    pub fn rama_state_to_host(s: *mut rama_session, buf: c_int, dst: *mut c_float, n: usize, n_out: *mut usize) -> c_int;
    pub fn rama_dev_alloc(ctx: *mut rama_ctx, n: usize, out: *mut *mut c_float) -> c_int;
    pub fn rama_dev_free(ctx: *mut rama_ctx, p: *mut c_float) -> c_int;
    pub fn rama_dev_h2d(ctx: *mut rama_ctx, dst: *mut c_float, src: *const c_float, n: usize) -> c_int;
    pub fn rama_dev_d2h(ctx: *mut rama_ctx, dst: *mut c_float, src: *const c_float, n: usize) -> c_int;
    pub fn rama_op_array_add(ctx: *mut rama_ctx, t: *mut c_float, s: *const c_float, n: usize) -> c_int;
    pub fn rama_op_array_mult(ctx: *mut rama_ctx, t: *mut c_float, s: *const c_float, n: usize) -> c_int;
    pub fn rama_op_sinu(ctx: *mut rama_ctx, o: *mut c_float, n: usize) -> c_int;
    pub fn rama_op_multi_head_attention(ctx: *mut rama_ctx, xb: *mut c_float, att: *mut c_float, q: *const c_float,
        key_cache: *const c_float, value_cache: *const c_float, cfg: *const rama_config, layer: i32, pos: i32) -> c_int;
    pub fn rama_op_copy_from_slice(ctx: *mut rama_ctx, t: *mut c_float, s: *const c_float, n: usize) -> c_int;
    pub fn rama_op_rmsnorm(ctx: *mut rama_ctx, o: *mut c_float, x: *const c_float, w: *const c_float, n: usize) -> c_int;
    pub fn rama_op_apply_position(ctx: *mut rama_ctx, q: *mut c_float, k: *mut c_float, pr: *const c_float,
        pi: *const c_float, head_size: usize) -> c_int;
    pub fn rama_op_matmul(ctx: *mut rama_ctx, o: *mut c_float, a: *const c_float, b: *const c_float, width: usize,
        o_rows: usize, o_cols: usize) -> c_int;
    pub fn rama_op_softmax(ctx: *mut rama_ctx, x: *mut c_float, n: usize) -> c_int;
    pub fn rama_op_sample(ctx: *mut rama_ctx, logits: *mut c_float, vocab: usize, temperature: c_float,
        topp: c_float, next: *mut i32) -> c_int;
    // no counterpart in the reference: prompt prefill and multi-request batching on tensor cores
    pub fn rama_prefill(s: *mut rama_session, tokens: *const i32, n: i32, pos0: i32, elapsed_ms: *mut c_float,
        ms_kind: *mut c_float, n_launch: *mut i32) -> c_int;
    pub fn rama_session_set_prefill(s: *mut rama_session, min_rows: i32) -> c_int;
    pub fn rama_batch_create(ctx: *mut rama_ctx, max_seqs: i32, out: *mut *mut rama_batch) -> c_int;
    pub fn rama_batch_destroy(b: *mut rama_batch) -> c_int;
    pub fn rama_forward_batch(b: *mut rama_batch, sessions: *const *mut rama_session, tokens: *const i32,
        pos: *const i32, n: i32) -> c_int;
    pub fn rama_sample_batch(b: *mut rama_batch, sessions: *const *mut rama_session, n: i32, temperature: c_float,
        topp: c_float, next: *mut i32) -> c_int;
    pub fn rama_batch_sync(b: *mut rama_batch) -> c_int;
}

#[repr(C)]
pub struct rama_batch { _private: [u8; 0] }

/// Reference behaviour on any device error is a panic (`.unwrap()` on every cudarc call, gpu.rs:73-209).
pub fn ck(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(rama_last_error()) }.to_string_lossy().into_owned();
        panic!("rama_b200 error {rc}: {msg}");
    }
}
