//! UNVERIFIED BY A COMPILER.  Deterministic getrandom backend.
//!
//! The reference draws every subset mask with `getrandom::fill` inside the private `CipheredBit::part`
//! (src/cipher.rs:92-97) and key material in `Polynomial::random` (src/polynomial.rs:87), so its ciphertexts cannot be
//! reproduced from outside.  getrandom 0.3+/0.4 lets the final binary replace the OS source: build with
//! `RUSTFLAGS='--cfg getrandom_backend="custom"'` and link this module, and every `getrandom::fill` of the process — the
//! reference's included — reads from ONE replayable stream.  The stream is the engine's documented Philox4x32-10 mask
//! stream (`hm_masks_generate_host`, include/hmgpu.h): bit-ciphertext u of the run takes mask bytes
//! [u * ceil(tau/8), (u+1) * ceil(tau/8)), so
//!
//!   homomorph::Context::encrypt(&v)            (reference, masks drawn through this backend, value by value)
//!   homomorph_gpu::Context::encrypt_with_masks(&[v..], &stream)   (engine, the same bytes)
//!
//! produce the same polynomials, and the decrypted results and canonical exports can be compared offline
//! (`hm_batch_download_canonical`).  FOR TESTS ONLY: Philox keyed by a 64-bit seed is not a CSPRNG.
use crate::ffi;
use std::sync::Mutex;

struct Stream {
    seed: u64,
    tau: u16,
    bytes: Vec<u8>, // the part of the stream generated so far
    pos: usize,
}

static STREAM: Mutex<Option<Stream>> = Mutex::new(None);

/// Arms the backend: from now on `getrandom::fill` returns the mask stream of (`seed`, `tau`) from its beginning.
pub fn seed_mask_stream(seed: u64, tau: u16) {
    *STREAM.lock().unwrap() = Some(Stream { seed, tau, bytes: Vec::new(), pos: 0 });
}

/// The first `units` masks of the stream, as the engine's `encrypt_with_masks` wants them.
pub fn mask_stream(seed: u64, tau: u16, units: usize) -> Vec<u8> {
    let mut out = vec![0u8; units * usize::from(tau).div_ceil(8)];
    let rc = unsafe { ffi::hm_masks_generate_host(tau, units, seed, out.as_mut_ptr()) };
    assert_eq!(rc, ffi::HM_OK);
    out
}

fn fill(dest: &mut [u8]) -> Result<(), getrandom::Error> {
    let mut guard = STREAM.lock().unwrap();
    let s = guard.as_mut().ok_or(getrandom::Error::UNSUPPORTED)?;
    let mb = usize::from(s.tau).div_ceil(8);
    while s.bytes.len() < s.pos + dest.len() {
        // extend by whole masks; the stream is a pure function of (seed, unit index), so regenerate the prefix
        let units = (s.pos + dest.len()).div_ceil(mb).max(2 * s.bytes.len() / mb).max(1024);
        s.bytes = mask_stream(s.seed, s.tau, units);
    }
    dest.copy_from_slice(&s.bytes[s.pos..s.pos + dest.len()]);
    s.pos += dest.len();
    Ok(())
}

/// getrandom's custom-backend entry point (getrandom 0.3 / 0.4: `getrandom_backend = "custom"`).
///
/// # Safety
/// `dest` must be valid for `len` writable bytes (guaranteed by getrandom).
#[no_mangle]
unsafe extern "Rust" fn __getrandom_v03_custom(dest: *mut u8, len: usize) -> Result<(), getrandom::Error> {
    fill(core::slice::from_raw_parts_mut(dest, len))
}
