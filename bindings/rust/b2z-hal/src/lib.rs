//! Safe wrapper of `b2z-sys` -- the Rust side of the drop-in boundary (uncompiled in this repository: the build image has
//! no Rust toolchain; `tests/test_rust_bindings.py` checks `b2z-sys` against `include/b2z.h` symbol by symbol).
//!
//! Layering, bottom up:
//!   * [`Buffer`]   -- device memory owned by Rust, freed on drop (risc0-zkp `hal::Buffer<T>`: `as_device_ptr`, `view`, `copy_from`).
//!   * [`B200Hal`]  -- one context on one GPU; methods carry the names and argument meaning of `trait Hal`
//!                     (risc0-zkp 3.0.2 `hal/mod.rs`) and `CircuitHal` (risc0-circuit-rv32im 4.0.2), so that
//!                     `impl Hal for B200Hal` is a one-line forward per method once the crate sources are reachable.
//!   * [`Circuit`]  -- a loaded circuit (taps + constraint program as a blob; generated kernels live inside libb2z).
//!   * [`Session`]  -- `ProverServer::prove_session`: segments in, ordered `SegmentReceipt`s out; [`Receipt::verify`]
//!                     is what zeth calls at /root/reference crates/host/src/bin/cli.rs:103.
//!
//! Threading: a context is single-threaded but not thread-bound (zeth builds its prover inside `spawn_blocking`,
//! /root/reference crates/host/src/lib.rs:131); every FFI entry point re-asserts the CUDA device.  `B200Hal` is `Send`
//! but not `Sync`, like upstream's `Rc<dyn Prover>` usage.
use anyhow::{anyhow, Result};
use b2z_sys as sys;
use std::ffi::CStr;
use std::os::raw::{c_char, c_void};
use std::ptr;

/// NULL = success; otherwise a library-owned string (valid until this thread's next failing call): copied, never freed.
fn chk(err: *const c_char) -> Result<()> {
    if err.is_null() {
        Ok(())
    } else {
        Err(anyhow!(unsafe { CStr::from_ptr(err) }.to_string_lossy().into_owned()))
    }
}

pub const INV_RATE: usize = 4;
pub const QUERIES: usize = 50;
pub const CHECK_SIZE: usize = 16;
pub const DIGEST_WORDS: usize = 8;

/// Device buffer of `len` u32 words (field elements are Montgomery words, ext elements 4 words, digests 8).
pub struct Buffer<'h> {
    hal: &'h B200Hal,
    ptr: *mut u32,
    len: usize,
}
impl<'h> Buffer<'h> {
    pub fn as_device_ptr(&self) -> *mut u32 {
        self.ptr
    }
    pub fn len(&self) -> usize {
        self.len
    }
    pub fn is_empty(&self) -> bool {
        self.len == 0
    }
    /// `Buffer::view` / `to_vec`: synchronising copy to the host.
    pub fn to_vec(&self) -> Result<Vec<u32>> {
        let mut v = vec![0u32; self.len];
        chk(unsafe { sys::b2z_d2h(self.hal.ctx, v.as_mut_ptr() as *mut c_void, self.ptr as *const c_void, 4 * self.len) })?;
        Ok(v)
    }
}
impl Drop for Buffer<'_> {
    fn drop(&mut self) {
        unsafe { sys::b2z_free(self.hal.ctx, self.ptr as *mut c_void) };
    }
}

/// One context on one B200 (upstream: `CudaHal` + the circuit hal).
pub struct B200Hal {
    ctx: *mut sys::b2z_ctx,
}
unsafe impl Send for B200Hal {}
impl B200Hal {
    pub fn new(device: i32) -> Result<Self> {
        let mut ctx = ptr::null_mut();
        chk(unsafe { sys::b2z_init(device, &mut ctx) })?;
        Ok(Self { ctx })
    }
    pub fn alloc(&self, words: usize) -> Result<Buffer<'_>> {
        let mut p: *mut c_void = ptr::null_mut();
        chk(unsafe { sys::b2z_alloc(self.ctx, 4 * words, &mut p) })?;
        Ok(Buffer { hal: self, ptr: p as *mut u32, len: words })
    }
    pub fn copy_from(&self, host: &[u32]) -> Result<Buffer<'_>> {
        let mut b = self.alloc(host.len().max(1))?;
        chk(unsafe { sys::b2z_h2d(self.ctx, b.ptr as *mut c_void, host.as_ptr() as *const c_void, 4 * host.len()) })?;
        chk(unsafe { sys::b2z_sync(self.ctx) })?;
        b.len = host.len(); // (a `Drop` type cannot be rebuilt with struct-update syntax)
        Ok(b)
    }
    // ---- trait Hal (risc0-zkp hal/mod.rs) ----
    pub fn batch_interpolate_ntt(&self, io: &Buffer, count: usize, po2: u32) -> Result<()> {
        chk(unsafe { sys::b2z_batch_interpolate_ntt(self.ctx, io.ptr, count, po2) })
    }
    pub fn batch_evaluate_ntt(&self, io: &Buffer, count: usize, po2: u32, expand_bits: u32) -> Result<()> {
        chk(unsafe { sys::b2z_batch_evaluate_ntt(self.ctx, io.ptr, count, po2, expand_bits) })
    }
    pub fn batch_expand_into_evaluate_ntt(&self, out: &Buffer, input: &Buffer, count: usize, in_po2: u32, expand_bits: u32) -> Result<()> {
        chk(unsafe { sys::b2z_batch_expand_into_evaluate_ntt(self.ctx, out.ptr, input.ptr, count, in_po2, expand_bits) })
    }
    pub fn batch_bit_reverse(&self, io: &Buffer, count: usize, po2: u32) -> Result<()> {
        chk(unsafe { sys::b2z_batch_bit_reverse(self.ctx, io.ptr, count, po2) })
    }
    pub fn zk_shift(&self, io: &Buffer, count: usize, po2: u32) -> Result<()> {
        chk(unsafe { sys::b2z_zk_shift(self.ctx, io.ptr, count, po2) })
    }
    pub fn hash_rows(&self, output: &Buffer, matrix: &Buffer, rows: usize, cols: usize) -> Result<()> {
        chk(unsafe { sys::b2z_hash_rows(self.ctx, output.ptr, matrix.ptr, rows, cols) })
    }
    pub fn hash_rows_chunk(&self, output: &Buffer, chunk: &Buffer, rows: usize, cols: usize, carry_in: bool, last: bool) -> Result<()> {
        chk(unsafe { sys::b2z_hash_rows_chunk(self.ctx, output.ptr, chunk.ptr, rows, cols, carry_in as i32, last as i32) })
    }
    pub fn hash_fold(&self, io: &Buffer, input_size: usize, output_size: usize) -> Result<()> {
        chk(unsafe { sys::b2z_hash_fold(self.ctx, io.ptr, input_size, output_size) })
    }
    pub fn merkle_build(&self, nodes: &Buffer, matrix: &Buffer, rows: usize, cols: usize) -> Result<()> {
        chk(unsafe { sys::b2z_merkle_build(self.ctx, nodes.ptr, matrix.ptr, rows, cols) })
    }
    pub fn fri_fold(&self, output: &Buffer, input: &Buffer, mix: &[u32; 4], out_count: usize) -> Result<()> {
        chk(unsafe { sys::b2z_fri_fold(self.ctx, output.ptr, input.ptr, mix.as_ptr(), out_count) })
    }
    pub fn batch_evaluate_any(&self, coeffs: &Buffer, poly_count: usize, po2: u32, which: &Buffer, xs: &Buffer, out: &Buffer) -> Result<()> {
        chk(unsafe { sys::b2z_batch_evaluate_any(self.ctx, coeffs.ptr, poly_count, po2, which.ptr, xs.ptr, which.len, out.ptr) })
    }
    pub fn mix_poly_coeffs(&self, output: &Buffer, mix_start: &[u32; 4], mix: &[u32; 4], input: &Buffer, combos: &Buffer, input_size: usize, count: usize) -> Result<()> {
        chk(unsafe { sys::b2z_mix_poly_coeffs(self.ctx, output.ptr, mix_start.as_ptr(), mix.as_ptr(), input.ptr, combos.ptr, input_size, count) })
    }
    pub fn eltwise_add_elem(&self, output: &Buffer, a: &Buffer, b: &Buffer) -> Result<()> {
        chk(unsafe { sys::b2z_eltwise_add_elem(self.ctx, output.ptr, a.ptr, b.ptr, output.len) })
    }
    pub fn eltwise_sum_extelem(&self, output: &Buffer, input: &Buffer, count: usize, to_add: usize) -> Result<()> {
        chk(unsafe { sys::b2z_eltwise_sum_extelem(self.ctx, output.ptr, input.ptr, count, to_add) })
    }
    pub fn eltwise_copy_elem(&self, output: &Buffer, input: &Buffer) -> Result<()> {
        chk(unsafe { sys::b2z_eltwise_copy_elem(self.ctx, output.ptr, input.ptr, output.len) })
    }
    pub fn eltwise_zeroize_elem(&self, io: &Buffer) -> Result<()> {
        chk(unsafe { sys::b2z_eltwise_zeroize_elem(self.ctx, io.ptr, io.len) })
    }
    pub fn gather_sample(&self, dst: &Buffer, src: &Buffer, idx: usize, size: usize, stride: usize) -> Result<()> {
        chk(unsafe { sys::b2z_gather_sample(self.ctx, dst.ptr, src.ptr, idx, size, stride) })
    }
    pub fn scatter(&self, into: &Buffer, index: &Buffer, offsets: &Buffer, values: &Buffer) -> Result<()> {
        chk(unsafe { sys::b2z_scatter(self.ctx, into.ptr, index.ptr, offsets.ptr, values.ptr, index.len.saturating_sub(1)) })
    }
    pub fn prefix_products(&self, io: &Buffer) -> Result<()> {
        chk(unsafe { sys::b2z_prefix_products(self.ctx, io.ptr, io.len / 4) })
    }
    pub fn poly_divide(&self, ext_coeffs: &Buffer, z: &[u32; 4]) -> Result<[u32; 4]> {
        let mut rem = [0u32; 4];
        chk(unsafe { sys::b2z_poly_divide(self.ctx, ext_coeffs.ptr, ext_coeffs.len / 4, z.as_ptr(), rem.as_mut_ptr()) })?;
        Ok(rem)
    }
    // ---- CircuitHal (risc0-circuit-rv32im) ----
    pub fn eval_check(&self, circuit: &Circuit, check: &Buffer, groups: [&Buffer; 3], globals: &Buffer, mix: &Buffer, poly_mix: &[u32; 4], po2: u32) -> Result<()> {
        let g = [groups[0].ptr as *const u32, groups[1].ptr as *const u32, groups[2].ptr as *const u32];
        chk(unsafe { sys::b2z_eval_check(self.ctx, circuit.h, check.ptr, g.as_ptr(), globals.ptr, mix.ptr, poly_mix.as_ptr(), po2) })
    }
    pub fn accumulate(&self, circuit: &Circuit, po2: u32, accum: &Buffer, code: &Buffer, data: &Buffer, mix: &Buffer, noise: &Buffer, zk_rows: usize) -> Result<()> {
        chk(unsafe { sys::b2z_accumulate(self.ctx, circuit.h, po2, accum.ptr, code.ptr, data.ptr, mix.ptr, noise.ptr, zk_rows) })
    }
    /// `SegmentProver::prove` on host buffers (pageable `Vec<u32>` memory is staged through the library's pinned ring).
    pub fn prove_segment(&self, circuit: &Circuit, po2: u32, code: &[u32], data: &[u32], globals: &[u32], accum_noise: &[u32], zk_rows: usize) -> Result<Vec<u32>> {
        let mut seal = vec![0u32; circuit.seal_words(po2)];
        let mut n = 0usize;
        chk(unsafe {
            sys::b2z_prove_segment(self.ctx, circuit.h, po2, code.as_ptr(), data.as_ptr(), globals.as_ptr(), accum_noise.as_ptr(), zk_rows, seal.as_mut_ptr(), &mut n)
        })?;
        seal.truncate(n);
        Ok(seal)
    }
    /// The control root (risc0: control ID) of a code group at `po2`.
    pub fn compute_control_root(&self, po2: u32, code: &Buffer, code_cols: usize) -> Result<[u32; 8]> {
        let mut root = [0u32; 8];
        chk(unsafe { sys::b2z_compute_control_root(self.ctx, po2, code.ptr, code_cols, root.as_mut_ptr()) })?;
        Ok(root)
    }
}
impl Drop for B200Hal {
    fn drop(&mut self) {
        unsafe { sys::b2z_destroy(self.ctx) };
    }
}

/// A circuit: taps + `PolyExtStep` program (risc0-zkp taps.rs / adapter.rs) as a blob.  `hal = None` loads it for
/// host-side verification only.
pub struct Circuit {
    h: *mut sys::b2z_circuit,
}
impl Circuit {
    pub fn load(hal: Option<&B200Hal>, blob: &[u32]) -> Result<Self> {
        let mut h = ptr::null_mut();
        chk(unsafe { sys::b2z_circuit_load(hal.map_or(ptr::null_mut(), |x| x.ctx), blob.as_ptr(), blob.len(), &mut h) })?;
        Ok(Self { h })
    }
    pub fn seal_words(&self, po2: u32) -> usize {
        unsafe { sys::b2z_circuit_seal_words(self.h, po2) }
    }
    pub fn set_control_root(&mut self, po2: u32, root: &[u32; 8]) -> Result<()> {
        chk(unsafe { sys::b2z_circuit_set_control_root(self.h, po2, root.as_ptr()) })
    }
    /// Per-segment STARK verifier (host CPU): Ok(()) or the rejection code.
    pub fn verify_segment(&self, seal: &[u32]) -> std::result::Result<(), i32> {
        match unsafe { sys::b2z_verify_segment(self.h, seal.as_ptr(), seal.len()) } {
            0 => Ok(()),
            rc => Err(rc),
        }
    }
}
impl Drop for Circuit {
    fn drop(&mut self) {
        unsafe { sys::b2z_circuit_free(self.h) };
    }
}

/// One segment's cached witness (the noise rows are an INPUT: SURVEY.md R2).
pub struct SegmentWitness<'a> {
    pub po2: u32,
    pub code: &'a [u32],
    pub data: &'a [u32],
    pub globals: &'a [u32],
    pub accum_noise: &'a [u32],
    pub zk_rows: usize,
}

/// `Receipt{inner: Composite{segments}, journal}` owned by the library; accessors copy out.
pub struct Receipt {
    r: *mut sys::b2z_receipt,
}
impl Receipt {
    pub fn journal(&self) -> &[u8] {
        let r = unsafe { &*self.r };
        if r.journal.is_null() { &[] } else { unsafe { std::slice::from_raw_parts(r.journal, r.journal_len) } }
    }
    pub fn segments(&self) -> &[sys::b2z_segment_receipt] {
        let r = unsafe { &*self.r };
        unsafe { std::slice::from_raw_parts(r.segments, r.n_segments) }
    }
    pub fn seal(&self, i: usize) -> &[u32] {
        let s = &self.segments()[i];
        unsafe { std::slice::from_raw_parts(s.seal, s.seal_words) }
    }
    /// `Receipt::verify(image_id)` -- /root/reference crates/host/src/bin/cli.rs:103.
    pub fn verify(&self, circuit: &Circuit, image_id: Option<&[u32; 8]>) -> Result<()> {
        let mut bad = 0usize;
        let rc = unsafe { sys::b2z_verify_receipt(circuit.h, self.r, image_id.map_or(ptr::null(), |d| d.as_ptr()), &mut bad) };
        if rc == 0 { Ok(()) } else { Err(anyhow!("receipt rejected: code {rc} at segment {bad}")) }
    }
    pub fn to_bytes(&self) -> Result<Vec<u8>> {
        let (mut p, mut n) = (ptr::null_mut(), 0usize);
        chk(unsafe { sys::b2z_receipt_serialize(self.r, &mut p, &mut n) })?;
        let v = unsafe { std::slice::from_raw_parts(p, n) }.to_vec();
        unsafe { sys::b2z_bytes_free(p) };
        Ok(v)
    }
    pub fn from_bytes(bytes: &[u8]) -> Result<Self> {
        let mut r = ptr::null_mut();
        chk(unsafe { sys::b2z_receipt_deserialize(bytes.as_ptr(), bytes.len(), &mut r) })?;
        Ok(Self { r })
    }
}
impl Drop for Receipt {
    fn drop(&mut self) {
        unsafe { sys::b2z_receipt_free(self.r) };
    }
}

/// `ProverServer::prove_session` on one GPU with `in_flight` segments proved concurrently.  What a
/// `impl risc0_zkvm::Prover for B200Prover` calls after the executor has produced the session's segments:
/// `prove_with_ctx(env, ctx, elf, opts)` = execute -> `Session` -> witnesses -> `prove_session` -> `Receipt`.
pub struct Session {
    s: *mut sys::b2z_session,
}
unsafe impl Send for Session {}
impl Session {
    pub fn new(device: i32, circuit_blob: &[u32], in_flight: u32) -> Result<Self> {
        let mut s = ptr::null_mut();
        chk(unsafe { sys::b2z_session_create(device, circuit_blob.as_ptr(), circuit_blob.len(), in_flight, &mut s) })?;
        Ok(Self { s })
    }
    pub fn set_control_root(&mut self, po2: u32, root: &[u32; 8]) -> Result<()> {
        chk(unsafe { sys::b2z_session_set_control_root(self.s, po2, root.as_ptr()) })
    }
    pub fn prove_session(&mut self, segments: &[SegmentWitness], journal: &[u8]) -> Result<Receipt> {
        let inputs: Vec<sys::b2z_segment_input> = segments
            .iter()
            .enumerate()
            .map(|(i, w)| sys::b2z_segment_input {
                index: i as u32,
                po2: w.po2,
                code: w.code.as_ptr(),
                data: w.data.as_ptr(),
                globals: w.globals.as_ptr(),
                accum_noise: w.accum_noise.as_ptr(),
                zk_rows: w.zk_rows,
                on_device: 0,
            })
            .collect();
        let mut r = ptr::null_mut();
        chk(unsafe { sys::b2z_prove_session(self.s, inputs.as_ptr(), inputs.len(), journal.as_ptr(), journal.len(), &mut r) })?;
        Ok(Receipt { r })
    }
}
impl Drop for Session {
    fn drop(&mut self) {
        unsafe { sys::b2z_session_destroy(self.s) };
    }
}
