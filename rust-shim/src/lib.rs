//! UNVERIFIED (never compiled): batched GPU counterpart of `homomorph::Context` / `Ciphered<T>` over libhmgpu.so.
//!
//! Same names and error types as the reference (src/context.rs:301-596, src/cipher.rs:126-259,
//! src/operations.rs); the difference is that a `CipheredBatch<T>` holds n values in HBM and every call
//! processes the whole batch.  Subset masks are an explicit argument (or a seed for the device-side Philox
//! stream) because `CipheredBit::part` (src/cipher.rs:92-97) draws from getrandom and cannot be reproduced.
pub mod ffi;

use core::marker::PhantomData;
use homomorph::prelude::*;

pub trait GpuOp: OperationRequirement {
    const CODE: i32;
}
impl GpuOp for HomomorphicAndGate { const CODE: i32 = 0; }
impl GpuOp for HomomorphicOrGate { const CODE: i32 = 1; }
impl GpuOp for HomomorphicXorGate { const CODE: i32 = 2; }
impl GpuOp for HomomorphicNotGate { const CODE: i32 = 3; }
impl GpuOp for HomomorphicAddition { const CODE: i32 = 4; }
impl GpuOp for HomomorphicMultiplication { const CODE: i32 = 5; }

#[derive(Debug)]
pub enum GpuError {
    Crypto(ContextCryptoError),
    Operation(OperationError),
    Cipher(CipherError),
    Engine { status: i32, detail: String },
}

/// Integer types whose bincode (fixint, little endian — src/cipher.rs:6-13) encoding is `to_le_bytes`.
pub trait Plain: Copy {
    const BITS: u32;
}
macro_rules! plain { ($($t:ty),+) => { $(impl Plain for $t { const BITS: u32 = <$t>::BITS; })+ } }
plain!(u8, u16, u32, u64, u128, i8, i16, i32, i64, i128);

pub struct GpuContext {
    raw: *mut ffi::hm_context,
    parameters: Parameters,
}

pub struct CipheredBatch<T: Plain> {
    raw: *mut ffi::hm_batch,
    ctx: *mut ffi::hm_context,
    _t: PhantomData<T>,
}

impl<T: Plain> Drop for CipheredBatch<T> {
    fn drop(&mut self) {
        unsafe { ffi::hm_batch_free(self.ctx, self.raw) }
    }
}
impl<T: Plain> CipheredBatch<T> {
    pub fn len(&self) -> usize { unsafe { ffi::hm_batch_len(self.raw) } }
    pub fn is_empty(&self) -> bool { self.len() == 0 }
}

impl Drop for GpuContext {
    fn drop(&mut self) {
        unsafe { ffi::hm_context_destroy(self.raw) }
    }
}

impl GpuContext {
    fn check(&self, rc: i32, op_req: Option<u16>) -> Result<(), GpuError> {
        match rc {
            ffi::HM_OK => Ok(()),
            ffi::HM_ERR_PUBLIC_KEY_UNSET => Err(GpuError::Crypto(ContextCryptoError::PublicKeyUnset)),
            ffi::HM_ERR_SECRET_KEY_UNSET => Err(GpuError::Crypto(ContextCryptoError::SecretKeyUnset)),
            ffi::HM_ERR_OPERATION_REQUIREMENT => Err(GpuError::Operation(OperationError::InvalidParameters {
                required_min_d_over_delta: op_req.unwrap_or(0),
                actual_d: self.parameters.d(),
                actual_delta: self.parameters.delta(),
            })),
            status => Err(GpuError::Engine { status, detail: String::new() }),
        }
    }

    pub fn new(parameters: Parameters, device: i32) -> Result<Self, GpuError> {
        let mut raw = core::ptr::null_mut();
        let rc = unsafe {
            ffi::hm_context_create(parameters.d(), parameters.dp(), parameters.delta(), parameters.tau(), device, &mut raw)
        };
        if rc != ffi::HM_OK {
            return Err(GpuError::Engine { status: rc, detail: "no usable CUDA device (no CPU fallback)".into() });
        }
        Ok(Self { raw, parameters })
    }

    /// `Context::set_secret_key` — also clears the public key (src/context.rs:568-571).
    pub fn set_secret_key(&mut self, sk: &SecretKey) -> Result<(), GpuError> {
        let bytes = sk.to_bytes();
        self.check(unsafe { ffi::hm_set_secret_key(self.raw, bytes.as_ptr(), bytes.len()) }, None)
    }

    pub fn set_public_key(&mut self, pk: &PublicKey) -> Result<(), GpuError> {
        let rows = pk.to_bytes();
        let ptrs: Vec<*const u8> = rows.iter().map(|r| r.as_ptr()).collect();
        let lens: Vec<usize> = rows.iter().map(|r| r.len()).collect();
        self.check(unsafe { ffi::hm_set_public_key(self.raw, ptrs.as_ptr(), lens.as_ptr(), rows.len()) }, None)
    }

    /// `Context::encrypt` for a slice; `masks` = n * T::BITS * ceil(tau/8) bytes (value-major, bit-minor).
    pub fn encrypt<T: Plain>(&self, values: &[T], masks: &[u8]) -> Result<CipheredBatch<T>, GpuError> {
        let mut out = core::ptr::null_mut();
        let rc = unsafe { ffi::hm_encrypt(self.raw, values.as_ptr().cast(), values.len(), T::BITS, masks.as_ptr(), &mut out) };
        self.check(rc, None)?;
        Ok(CipheredBatch { raw: out, ctx: self.raw, _t: PhantomData })
    }

    pub fn encrypt_seeded<T: Plain>(&self, values: &[T], seed: u64) -> Result<CipheredBatch<T>, GpuError> {
        let mut out = core::ptr::null_mut();
        let rc = unsafe { ffi::hm_encrypt_seeded(self.raw, values.as_ptr().cast(), values.len(), T::BITS, seed, &mut out) };
        self.check(rc, None)?;
        Ok(CipheredBatch { raw: out, ctx: self.raw, _t: PhantomData })
    }

    /// `Context::decrypt` (src/context.rs:480-488).
    pub fn decrypt<T: Plain + Default>(&self, c: &CipheredBatch<T>) -> Result<Vec<T>, GpuError> {
        let mut out = vec![T::default(); c.len()];
        self.check(unsafe { ffi::hm_decrypt(self.raw, c.raw, out.as_mut_ptr().cast()) }, None)?;
        Ok(out)
    }

    /// `Context::apply2` (src/context.rs:515-527): the requirement check happens inside hm_apply2.
    pub fn apply2<O: GpuOp, T: Plain>(&self, a: &CipheredBatch<T>, b: &CipheredBatch<T>) -> Result<CipheredBatch<T>, GpuError> {
        let mut out = core::ptr::null_mut();
        let rc = unsafe { ffi::hm_apply2(self.raw, O::CODE, a.raw, b.raw, &mut out) };
        self.check(rc, Some(O::MIN_D_OVER_DELTA))?;
        Ok(CipheredBatch { raw: out, ctx: self.raw, _t: PhantomData })
    }

    /// `Context::apply1` (src/context.rs:496-507): in place.
    pub fn apply1<O: GpuOp, T: Plain>(&self, a: &mut CipheredBatch<T>) -> Result<(), GpuError> {
        self.check(unsafe { ffi::hm_apply1(self.raw, O::CODE, a.raw) }, Some(O::MIN_D_OVER_DELTA))
    }
}
