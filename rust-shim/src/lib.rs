//! UNVERIFIED BY A COMPILER (this image has no Rust toolchain) — batched GPU back end for the `homomorph` crate over
//! libhmgpu.so.  `ffi.rs` is generated from include/hmgpu.h and checked against it by tests/test_rust_ffi.py.
//!
//! `Context` below has the method names, argument order and error enums of the reference's
//! `homomorph::Context` (src/context.rs:301-596): `new`, `parameters`, `generate_secret_key`, `generate_public_key`,
//! `get_secret_key`, `get_public_key`, `set_secret_key`, `set_public_key`, `encrypt`, `decrypt`, `apply1`, `apply2`.
//! The difference is the unit of work: `Ciphered<'ctx, T>` is a BATCH of n values resident in HBM
//! (src/cipher.rs:126-130 is one value on the host), `encrypt` takes a slice and `decrypt` returns a `Vec`.
//! The reference's key types, `Parameters`, error enums and operation markers are reused from the crate itself, so a
//! program written against `homomorph::prelude` changes its `use` line and the element type of its data.
//!
//! Randomness: the reference draws subset masks from `getrandom` inside a private function
//! (`CipheredBit::part`, src/cipher.rs:92-97), the engine takes them as bytes.  `encrypt` fills them with the same
//! `getrandom::fill` the reference uses; `mask_rng` holds a deterministic getrandom custom backend, so that the
//! unmodified reference crate and this back end can be driven by ONE mask stream and compared.
pub mod ffi;
pub mod mask_rng;

use core::marker::PhantomData;
use homomorph::prelude::*;

/// The engine's selector of an operation marker (src/impls/numbers.rs:7-50).
pub trait GpuOperation: OperationRequirement {
    const CODE: i32;
}
impl GpuOperation for HomomorphicAndGate { const CODE: i32 = ffi::HM_OP_AND; }
impl GpuOperation for HomomorphicOrGate { const CODE: i32 = ffi::HM_OP_OR; }
impl GpuOperation for HomomorphicXorGate { const CODE: i32 = ffi::HM_OP_XOR; }
impl GpuOperation for HomomorphicNotGate { const CODE: i32 = ffi::HM_OP_NOT; }
impl GpuOperation for HomomorphicAddition { const CODE: i32 = ffi::HM_OP_ADD; }
impl GpuOperation for HomomorphicMultiplication { const CODE: i32 = ffi::HM_OP_MUL; }

/// Failures that have no counterpart in the reference (it has no device): the status of include/hmgpu.h.
#[derive(Debug)]
pub struct EngineError {
    pub status: i32,
    pub detail: String,
}

/// What `Context::new` can report besides the reference's panics: no usable CUDA device (there is no CPU fallback).
#[derive(Debug)]
pub enum GpuError {
    Crypto(ContextCryptoError),
    Operation(OperationError),
    Engine(EngineError),
}

/// Integer types whose bincode encoding (fixint, little endian — src/cipher.rs:6-13, :176) is `to_le_bytes`, so a slice of
/// them on a little-endian host is already the byte stream `Ciphered::try_cipher` would produce value by value.
pub trait Plain: Copy + Default {
    const BITS: u32;
}
macro_rules! plain { ($($t:ty),+) => { $(impl Plain for $t { const BITS: u32 = <$t>::BITS; })+ } }
plain!(u8, u16, u32, u64, u128, i8, i16, i32, i64, i128);

#[cfg(not(target_endian = "little"))]
compile_error!("the slice -> byte stream shortcut of `Plain` needs a little-endian host");

pub struct Context {
    raw: *mut ffi::hm_context,
    parameters: Parameters,
    secret_key: Option<SecretKey>,
    public_key: Option<PublicKey>,
}

/// n values x `T::BITS` bit-ciphertexts in HBM.  Borrowing the context keeps the device memory valid: a batch cannot
/// outlive the `Context` that owns it.
pub struct Ciphered<'ctx, T: Plain> {
    raw: *mut ffi::hm_batch,
    _ctx: PhantomData<&'ctx Context>,
    _t: PhantomData<T>,
}

impl<T: Plain> Drop for Ciphered<'_, T> {
    fn drop(&mut self) {
        // the owning context is recorded in the batch (include/hmgpu.h, hm_batch_free)
        unsafe { ffi::hm_batch_free(core::ptr::null_mut(), self.raw) }
    }
}

impl<T: Plain> Ciphered<'_, T> {
    /// Number of values in the batch.
    pub fn len(&self) -> usize {
        unsafe { ffi::hm_batch_len(self.raw) }
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::hm_context_destroy(self.raw) } // zeroises the secret key and its tables (src/context.rs:199-206)
    }
}

impl Context {
    fn engine(&self, status: i32) -> EngineError {
        let detail = unsafe { core::ffi::CStr::from_ptr(ffi::hm_last_error(self.raw)) }.to_string_lossy().into_owned();
        EngineError { status, detail }
    }

    fn crypto(&self, rc: i32) -> Result<(), ContextCryptoError> {
        match rc {
            ffi::HM_OK => Ok(()),
            ffi::HM_ERR_PUBLIC_KEY_UNSET => Err(ContextCryptoError::PublicKeyUnset),
            ffi::HM_ERR_SECRET_KEY_UNSET => Err(ContextCryptoError::SecretKeyUnset),
            ffi::HM_ERR_INVALID_LENGTH => Err(ContextCryptoError::Cipher(CipherError::InvalidCipheredLength { len: 0 })),
            status => panic!("libhmgpu: {:?}", self.engine(status)),
        }
    }

    /// `Context::new` (src/context.rs:341-347) on CUDA device `device`.
    pub fn new(parameters: Parameters, device: i32) -> Result<Self, GpuError> {
        let mut raw = core::ptr::null_mut();
        let rc = unsafe {
            ffi::hm_context_create(parameters.d(), parameters.dp(), parameters.delta(), parameters.tau(), device, &mut raw)
        };
        if rc != ffi::HM_OK {
            return Err(GpuError::Engine(EngineError { status: rc, detail: "no usable CUDA device (the engine has no CPU fallback)".into() }));
        }
        Ok(Self { raw, parameters, secret_key: None, public_key: None })
    }

    pub const fn parameters(&self) -> &Parameters {
        &self.parameters
    }
    pub const fn get_secret_key(&self) -> Option<&SecretKey> {
        self.secret_key.as_ref()
    }
    pub const fn get_public_key(&self) -> Option<&PublicKey> {
        self.public_key.as_ref()
    }

    /// src/context.rs:421-425
    pub fn generate_secret_key(&mut self) {
        let sk = SecretKey::random(self.parameters.d());
        self.set_secret_key(sk);
    }

    /// src/context.rs:440-454
    pub fn generate_public_key(&mut self) -> Result<(), ContextCryptoError> {
        let sk = self.secret_key.as_ref().ok_or(ContextCryptoError::SecretKeyUnset)?;
        let pk = PublicKey::random(self.parameters.dp(), self.parameters.delta(), self.parameters.tau(), sk);
        self.set_public_key(pk);
        Ok(())
    }

    /// src/context.rs:568-571 — also clears the public key, on the device as well.
    pub fn set_secret_key(&mut self, secret_key: SecretKey) {
        let bytes = secret_key.to_bytes();
        let rc = unsafe { ffi::hm_set_secret_key(self.raw, bytes.as_ptr(), bytes.len()) };
        assert_eq!(rc, ffi::HM_OK, "libhmgpu: {:?}", self.engine(rc));
        self.secret_key = Some(secret_key);
        self.public_key = None;
    }

    /// src/context.rs:592-595
    pub fn set_public_key(&mut self, public_key: PublicKey) {
        let rows = public_key.to_bytes();
        let ptrs: Vec<*const u8> = rows.iter().map(|r| r.as_ptr()).collect();
        let lens: Vec<usize> = rows.iter().map(Vec::len).collect();
        let rc = unsafe { ffi::hm_set_public_key(self.raw, ptrs.as_ptr(), lens.as_ptr(), rows.len()) };
        assert_eq!(rc, ffi::HM_OK, "libhmgpu: {:?}", self.engine(rc));
        self.public_key = Some(public_key);
    }

    fn mask_len<T: Plain>(&self, n: usize) -> usize {
        n * T::BITS as usize * usize::from(self.parameters.tau()).div_ceil(8)
    }

    /// `Context::encrypt` (src/context.rs:463-471) for a slice of values.  The subset masks are drawn with the same
    /// `getrandom::fill` as `CipheredBit::part` (src/cipher.rs:92-97), n * T::BITS * ceil(tau/8) bytes in the order the
    /// reference would draw them (value by value, bit by bit).
    pub fn encrypt<T: Plain>(&self, data: &[T]) -> Result<Ciphered<'_, T>, ContextCryptoError> {
        let mut masks = vec![0u8; self.mask_len::<T>(data.len())];
        getrandom::fill(&mut masks).map_err(|_| ContextCryptoError::Cipher(CipherError::Randomness))?;
        self.encrypt_with_masks(data, &masks)
    }

    /// The same with caller-supplied masks (reproducible runs, parity tests).
    pub fn encrypt_with_masks<T: Plain>(&self, data: &[T], masks: &[u8]) -> Result<Ciphered<'_, T>, ContextCryptoError> {
        assert_eq!(masks.len(), self.mask_len::<T>(data.len()), "masks must hold n * T::BITS * ceil(tau/8) bytes");
        let mut out = core::ptr::null_mut();
        let rc = unsafe { ffi::hm_encrypt(self.raw, data.as_ptr().cast(), data.len(), T::BITS, masks.as_ptr(), &mut out) };
        self.crypto(rc)?;
        Ok(Ciphered { raw: out, _ctx: PhantomData, _t: PhantomData })
    }

    /// Masks generated on the device from a seed (Philox4x32-10): reproducibility and benchmarks only, not a CSPRNG.
    pub fn encrypt_seeded<T: Plain>(&self, data: &[T], seed: u64) -> Result<Ciphered<'_, T>, ContextCryptoError> {
        let mut out = core::ptr::null_mut();
        let rc = unsafe { ffi::hm_encrypt_seeded(self.raw, data.as_ptr().cast(), data.len(), T::BITS, seed, &mut out) };
        self.crypto(rc)?;
        Ok(Ciphered { raw: out, _ctx: PhantomData, _t: PhantomData })
    }

    /// `Context::decrypt` (src/context.rs:480-488).
    pub fn decrypt<T: Plain>(&self, ciphered: &Ciphered<'_, T>) -> Result<Vec<T>, ContextCryptoError> {
        let mut out = vec![T::default(); ciphered.len()];
        self.crypto(unsafe { ffi::hm_decrypt(self.raw, ciphered.raw, out.as_mut_ptr().cast()) })?;
        Ok(out)
    }

    fn operation<O: GpuOperation>(&self, rc: i32) -> Result<(), OperationError> {
        match rc {
            ffi::HM_OK => Ok(()),
            ffi::HM_ERR_OPERATION_REQUIREMENT => Err(OperationError::InvalidParameters {
                required_min_d_over_delta: O::MIN_D_OVER_DELTA,
                actual_d: self.parameters.d(),
                actual_delta: self.parameters.delta(),
            }),
            status => panic!("libhmgpu: {:?}", self.engine(status)),
        }
    }

    /// `Context::apply1` (src/context.rs:496-507): in place; the requirement check happens inside hm_apply1.
    pub fn apply1<O: GpuOperation, T: Plain>(&self, a: &mut Ciphered<'_, T>) -> Result<(), OperationError> {
        self.operation::<O>(unsafe { ffi::hm_apply1(self.raw, O::CODE, a.raw) })
    }

    /// `Context::apply2` (src/context.rs:515-527).
    pub fn apply2<'ctx, O: GpuOperation, T: Plain>(&'ctx self, a: &Ciphered<'ctx, T>, b: &Ciphered<'ctx, T>) -> Result<Ciphered<'ctx, T>, OperationError> {
        let mut out = core::ptr::null_mut();
        self.operation::<O>(unsafe { ffi::hm_apply2(self.raw, O::CODE, a.raw, b.raw, &mut out) })?;
        Ok(Ciphered { raw: out, _ctx: PhantomData, _t: PhantomData })
    }

    /// The reference's `unsafe { O::apply(a, b) }` (src/operations.rs:140): no parameter check.
    ///
    /// # Safety
    /// As in the reference: the caller asserts that the parameters support the operation.
    pub unsafe fn apply2_unchecked<'ctx, O: GpuOperation, T: Plain>(&'ctx self, a: &Ciphered<'ctx, T>, b: &Ciphered<'ctx, T>) -> Ciphered<'ctx, T> {
        let mut out = core::ptr::null_mut();
        let rc = ffi::hm_apply2_unchecked(self.raw, O::CODE, a.raw, b.raw, &mut out);
        assert_eq!(rc, ffi::HM_OK, "libhmgpu: {:?}", self.engine(rc));
        Ciphered { raw: out, _ctx: PhantomData, _t: PhantomData }
    }

    /// Waits for everything enqueued on this context's stream (operations are asynchronous, results stay in HBM).
    pub fn synchronize(&self) {
        let rc = unsafe { ffi::hm_context_synchronize(self.raw) };
        assert_eq!(rc, ffi::HM_OK, "libhmgpu: {:?}", self.engine(rc));
    }
}
