//! FFI declarations of include/mapquik_b200.h plus a thin safe wrapper shaped like the reference's own types
//! (`Index` / `ReadOnlyIndex`, src/index.rs) so that src/closures.rs changes by a few lines only.
//! NOT compiled in the build image (no Rust toolchain); kept next to the C ABI it mirrors.
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct MqParams { pub k: u32, pub l: u32, pub density: f64, pub use_hpc: u32, pub c: u32, pub s: u32, pub g: u32 }

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct MqHit {
    pub mapped: u8, pub rc: u8, pub mapq: u8, pub pad_: u8, pub ref_idx: u32,
    pub q_start: u64, pub q_end: u64, pub r_start: u64, pub r_end: u64, pub score: u64,
}

/// one interval of bytes other than A/C/G/T inside a packed sequence array (include/mapquik_b200.h `mq_exc`)
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct MqExc { pub start: u64, pub len: u32, pub byte: u32 }

/// packed sequence array (`mq_packed`): 2-bit codes, block bitmap, exception intervals
#[repr(C)]
pub struct MqPacked { pub words: *const u32, pub flags: *const u32, pub exc: *const MqExc, pub n_exc: u64, pub n_bases: u64 }

pub enum MqCtx {}

extern "C" {
    pub fn mq_create(out: *mut *mut MqCtx, p: *const MqParams, device: c_int) -> c_int;
    pub fn mq_create_multi(out: *mut *mut MqCtx, p: *const MqParams, devices: *const c_int, n_devices: c_int) -> c_int;
    pub fn mq_set_host_threads(c: *mut MqCtx, n: c_int) -> c_int;
    pub fn mq_packed_words(n_bases: u64) -> u64;
    pub fn mq_packed_flag_words(n_bases: u64) -> u64;
    pub fn mq_pack_at(ascii: *const u8, n_bases: u64, at_base: u64, words: *mut u32, flags: *mut u32, exc: *mut MqExc, exc_cap: u64,
                      n_exc: *mut u64, fold_case: c_int) -> c_int;
    pub fn mq_index_add_packed(c: *mut MqCtx, seqs: *const MqPacked, offs: *const u64, n: u32, first_ref_idx: u32, nb_mers_out: *mut u64) -> c_int;
    pub fn mq_map_batch_packed(c: *mut MqCtx, seqs: *const MqPacked, offs: *const u64, n: u32, out: *mut MqHit) -> c_int;
    pub fn mq_destroy(c: *mut MqCtx);
    pub fn mq_strerror(code: c_int) -> *const c_char;
    pub fn mq_last_error(c: *const MqCtx) -> *const c_char;
    pub fn mq_index_add(c: *mut MqCtx, seqs: *const u8, offs: *const u64, n: u32, first_ref_idx: u32, nb_mers_out: *mut u64) -> c_int;
    pub fn mq_index_freeze(c: *mut MqCtx, ref_lens: *const u64, n_refs: u32, n_unique: *mut u64, n_keys: *mut u64) -> c_int;
    pub fn mq_map_batch(c: *mut MqCtx, seqs: *const u8, offs: *const u64, n: u32, out: *mut MqHit) -> c_int;
}

fn check(c: *const MqCtx, rc: c_int, what: &str) {
    if rc != 0 {
        // the reference aborts on every error (panic = "abort", Cargo.toml:49); keep that convention
        let msg = unsafe { CStr::from_ptr(mq_strerror(rc)) }.to_string_lossy().into_owned();
        let detail = if c.is_null() { String::new() } else { unsafe { CStr::from_ptr(mq_last_error(c)) }.to_string_lossy().into_owned() };
        panic!("{}: {} ({})", what, msg, detail);
    }
}

/// Stands in for `Index` (src/index.rs:73-105) while the reference is being read ...
pub struct GpuIndex { ctx: *mut MqCtx, ref_lens: Vec<u64> }
/// ... and for `ReadOnlyIndex` (src/index.rs:108-128) afterwards.
pub struct GpuReadOnlyIndex { ctx: *mut MqCtx }
unsafe impl Send for GpuIndex {}
unsafe impl Send for GpuReadOnlyIndex {}

impl GpuIndex {
    /// ≙ Index::new() (closures.rs:24); `params` carries k, l, density, use_hpc, c, s, g of `Params` (main.rs:33-47)
    pub fn new(params: MqParams, device: i32) -> Self {
        let mut ctx: *mut MqCtx = std::ptr::null_mut();
        check(std::ptr::null(), unsafe { mq_create(&mut ctx, &params, device) }, "mq_create");
        GpuIndex { ctx, ref_lens: Vec::new() }
    }
    /// The same over several GPUs of the box (closures.rs:85,183 fan out over threads; here over devices): the library
    /// partitions the reference by base range, exchanges the minimizer stores GPU to GPU at `freeze`, shards the reads.
    pub fn new_multi(params: MqParams, devices: &[i32]) -> Self {
        let mut ctx: *mut MqCtx = std::ptr::null_mut();
        check(std::ptr::null(), unsafe { mq_create_multi(&mut ctx, &params, devices.as_ptr(), devices.len() as c_int) }, "mq_create_multi");
        GpuIndex { ctx, ref_lens: Vec::new() }
    }
    /// `--threads` (main.rs:138-141, 212): host threads `find_matches_batch` may use to pack ASCII batches on the fly
    /// (sub-batches from the back of the batch are packed to 2 bits per base while ASCII ones cross PCIe from the front)
    pub fn set_host_threads(&mut self, threads: usize) {
        check(self.ctx, unsafe { mq_set_host_threads(self.ctx, threads as c_int) }, "mq_set_host_threads");
    }
    /// ≙ mers::ref_extract for a batch of upper-cased records (closures.rs:48,63); returns the k-min-mer count of each
    pub fn ref_extract_batch(&mut self, seqs: &[u8], offs: &[u64]) -> Vec<u64> {
        let n = (offs.len() - 1) as u32;
        let mut nb = vec![0u64; n as usize];
        let first = self.ref_lens.len() as u32;
        check(self.ctx, unsafe { mq_index_add(self.ctx, seqs.as_ptr(), offs.as_ptr(), n, first, nb.as_mut_ptr()) }, "mq_index_add");
        for w in offs.windows(2) { self.ref_lens.push(w[1] - w[0]); }
        nb
    }
    /// ≙ mers_index.get_count() + ReadOnlyIndex::new(mers_index.index) (closures.rs:92,94)
    pub fn freeze(self) -> (GpuReadOnlyIndex, u64) {
        let mut n_unique = 0u64;
        check(self.ctx, unsafe { mq_index_freeze(self.ctx, self.ref_lens.as_ptr(), self.ref_lens.len() as u32, &mut n_unique, std::ptr::null_mut()) },
              "mq_index_freeze");
        let ctx = self.ctx;
        std::mem::forget(self);
        (GpuReadOnlyIndex { ctx }, n_unique)
    }
}
impl Drop for GpuIndex { fn drop(&mut self) { unsafe { mq_destroy(self.ctx) } } }

impl GpuReadOnlyIndex {
    /// ≙ mers::find_matches for a batch of upper-cased reads (closures.rs:102,106): one MqHit per read, input order.
    /// The caller formats `q_id, q_len, q_start, q_end, ±, r_id, r_len, r_start, r_end, score, r_len, mapq` (mers.rs:181)
    /// for every hit with `mapped != 0`.
    pub fn find_matches_batch(&self, seqs: &[u8], offs: &[u64]) -> Vec<MqHit> {
        let n = (offs.len() - 1) as u32;
        let mut hits = vec![MqHit::default(); n as usize];
        check(self.ctx, unsafe { mq_map_batch(self.ctx, seqs.as_ptr(), offs.as_ptr(), n, hits.as_mut_ptr()) }, "mq_map_batch");
        hits
    }
}
impl Drop for GpuReadOnlyIndex { fn drop(&mut self) { unsafe { mq_destroy(self.ctx) } } }

/// A batch buffer the record closures append to INSTEAD of `record.seq().to_ascii_uppercase()` (closures.rs:63,106):
/// the one pass the reference spends on upper-casing a copy packs the record to 2 bits per base (upper-casing folded
/// in), so a quarter of the bytes cross PCIe.  `mq_pack_at` may be called concurrently for disjoint ranges, i.e. from
/// seq_io's worker closures once each record has been given its destination offset.
pub struct PackedBatch { pub words: Vec<u32>, pub flags: Vec<u32>, pub exc: Vec<MqExc>, pub offs: Vec<u64> }
impl PackedBatch {
    pub fn with_capacity(bases: u64) -> Self {
        let (w, f) = unsafe { (mq_packed_words(bases), mq_packed_flag_words(bases)) };
        PackedBatch { words: vec![0u32; w as usize], flags: vec![0u32; f as usize], exc: Vec::new(), offs: vec![0] }
    }
    pub fn push_record(&mut self, seq: &[u8]) {
        let at = *self.offs.last().unwrap();
        let mut tmp = vec![MqExc::default(); 64];
        let mut n = 0u64;
        let mut rc = unsafe { mq_pack_at(seq.as_ptr(), seq.len() as u64, at, self.words.as_mut_ptr(), self.flags.as_mut_ptr(), tmp.as_mut_ptr(), tmp.len() as u64, &mut n, 1) };
        if rc == -5 {   // MQ_ERR_RANGE: more intervals than the scratch list holds; the codes are in place, list them again
            tmp.resize(n as usize, MqExc::default());
            rc = unsafe { mq_pack_at(seq.as_ptr(), seq.len() as u64, at, self.words.as_mut_ptr(), self.flags.as_mut_ptr(), tmp.as_mut_ptr(), tmp.len() as u64, &mut n, 1) };
        }
        check(std::ptr::null(), rc, "mq_pack_at");
        self.exc.extend_from_slice(&tmp[..n as usize]);
        self.offs.push(at + seq.len() as u64);
    }
    fn view(&self) -> MqPacked {
        MqPacked { words: self.words.as_ptr(), flags: self.flags.as_ptr(), exc: self.exc.as_ptr(), n_exc: self.exc.len() as u64, n_bases: *self.offs.last().unwrap() }
    }
}
impl GpuReadOnlyIndex {
    /// ≙ mers::find_matches for a packed batch
    pub fn find_matches_packed(&self, b: &PackedBatch) -> Vec<MqHit> {
        let n = (b.offs.len() - 1) as u32;
        let mut hits = vec![MqHit::default(); n as usize];
        let v = b.view();
        check(self.ctx, unsafe { mq_map_batch_packed(self.ctx, &v, b.offs.as_ptr(), n, hits.as_mut_ptr()) }, "mq_map_batch_packed");
        hits
    }
}
