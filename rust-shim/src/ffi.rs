//! UNVERIFIED raw bindings of include/hmgpu.h (one line per declaration that the safe layer uses).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct hm_context {
    _p: [u8; 0],
}
#[repr(C)]
pub struct hm_batch {
    _p: [u8; 0],
}

pub const HM_OK: c_int = 0;
pub const HM_ERR_INVALID_PARAMETERS: c_int = -1;
pub const HM_ERR_PUBLIC_KEY_UNSET: c_int = -2;
pub const HM_ERR_SECRET_KEY_UNSET: c_int = -3;
pub const HM_ERR_OPERATION_REQUIREMENT: c_int = -4;
pub const HM_ERR_INVALID_LENGTH: c_int = -5;
pub const HM_ERR_CUDA: c_int = -6;

extern "C" {
    pub fn hm_status_string(status: c_int) -> *const c_char;
    pub fn hm_last_error(ctx: *const hm_context) -> *const c_char;
    pub fn hm_device_count() -> c_int;
    pub fn hm_context_create(d: u16, dp: u16, delta: u16, tau: u16, device: c_int, out: *mut *mut hm_context) -> c_int;
    pub fn hm_context_destroy(ctx: *mut hm_context);
    pub fn hm_set_secret_key(ctx: *mut hm_context, bytes: *const u8, len: usize) -> c_int;
    pub fn hm_set_public_key(ctx: *mut hm_context, polys: *const *const u8, lens: *const usize, n: usize) -> c_int;
    pub fn hm_encrypt(ctx: *mut hm_context, values: *const u8, n: usize, bits: u32, masks: *const u8, out: *mut *mut hm_batch) -> c_int;
    pub fn hm_encrypt_seeded(ctx: *mut hm_context, values: *const u8, n: usize, bits: u32, seed: u64, out: *mut *mut hm_batch) -> c_int;
    pub fn hm_masks_generate_host(tau: u16, units: usize, seed: u64, masks_out: *mut u8) -> c_int;
    pub fn hm_decrypt(ctx: *mut hm_context, b: *const hm_batch, values_out: *mut u8) -> c_int;
    pub fn hm_apply2(ctx: *mut hm_context, op: c_int, a: *const hm_batch, b: *const hm_batch, out: *mut *mut hm_batch) -> c_int;
    pub fn hm_apply1(ctx: *mut hm_context, op: c_int, a: *mut hm_batch) -> c_int;
    pub fn hm_batch_len(b: *const hm_batch) -> usize;
    pub fn hm_batch_bits(b: *const hm_batch) -> u32;
    pub fn hm_batch_value_words(b: *const hm_batch) -> usize;
    pub fn hm_batch_download(ctx: *mut hm_context, b: *const hm_batch, host: *mut u64) -> c_int;
    pub fn hm_batch_free(ctx: *mut hm_context, b: *mut hm_batch);
    pub fn hm_batch_slice(ctx: *mut hm_context, src: *const hm_batch, first_bit: u32, n_bits: u32, out: *mut *mut hm_batch) -> c_int;
    pub fn hm_batch_concat(ctx: *mut hm_context, parts: *const *const hm_batch, count: usize, out: *mut *mut hm_batch) -> c_int;
    pub fn hm_apply2_fields(ctx: *mut hm_context, op: c_int, a: *const hm_batch, b: *const hm_batch, field_bits: *const u32, n_fields: usize, out: *mut *mut hm_batch) -> c_int;
    pub fn hm_host_alloc(bytes: usize) -> *mut c_void;
    pub fn hm_host_free(p: *mut c_void);
}
