// NOT COMPILED in this repository's environment (no rustc/cargo in the image).
// Reference-side binding of libgenedex_b200.so, kept in sync with INTEGRATION.md and include/genedex_b200.h.
pub mod ffi;
use crate::{Cursor, FmIndex, HalfOpenInterval, Hit, IndexStorage, text_with_rank_support::TextWithRankSupport};

/// Device replica of an `FmIndex`. `FmIndex: Clone` clones the `Arc`, not device memory.
pub(crate) struct DeviceIndex(std::sync::Arc<Handle>);
struct Handle(*mut ffi::gdx_index);
unsafe impl Send for Handle {}      // the C ABI is re-entrant on one handle (header, "Conventions")
unsafe impl Sync for Handle {}
impl Drop for Handle { fn drop(&mut self) { unsafe { ffi::gdx_index_destroy(self.0) } } }

fn check(status: ffi::gdx_status) {
    match status {
        0 => {}
        // same message as src/alphabet.rs:197, so existing `#[should_panic]` tests keep passing
        1 => panic!("symbol in io representation should be valid (query {})",
                    unsafe { ffi::gdx_last_error_query() }),
        _ => panic!("genedex_b200: {}", unsafe {
            std::ffi::CStr::from_ptr(ffi::gdx_last_error_message()).to_string_lossy() }),
    }
}

/// Drain `impl IntoIterator<Item = Q: AsRef<[u8]>>` into (bytes, offsets): the only host-side work.
/// Plain `Vec`s are the intended form: for DNA alphabets the library packs the bytes to 2 bits per symbol
/// with its own thread pool while it stages them into pinned memory (a quarter of the PCIe bytes), sends the
/// offsets chunk-relative as u32 and brings counts back as u32 -- pinned caller buffers buy nothing there.
/// Callers that already hold 2-bit packed reads set `encoding: GDX_QUERIES_PACKED_2BIT` (ideally in memory
/// from `gdx_host_alloc`) and skip the packing stage.
fn pack<Q: AsRef<[u8]>>(queries: impl IntoIterator<Item = Q>) -> (Vec<u8>, Vec<u64>) {
    let (mut bytes, mut offsets) = (Vec::new(), vec![0u64]);
    for q in queries { bytes.extend_from_slice(q.as_ref()); offsets.push(bytes.len() as u64); }
    (bytes, offsets)
}

impl<I: IndexStorage, R: TextWithRankSupport<I>> FmIndex<I, R> {
    // replaces BatchComputedCursors::new(..) in src/lib.rs:241-246; same signature, same order
    pub fn cursors_for_many_queries<'a, Q: AsRef<[u8]>>(
        &'a self, queries: impl IntoIterator<Item = Q>,
    ) -> impl Iterator<Item = Cursor<'a, I, R>> {
        let (bytes, offsets) = pack(queries);
        let nq = offsets.len() - 1;
        let (mut starts, mut ends) = (vec![0u64; nq], vec![0u64; nq]);
        let q = ffi::gdx_queries { bytes: bytes.as_ptr(), offsets: offsets.as_ptr(), fixed_len: 0, nq: nq as u64,
                                   encoding: ffi::GDX_QUERIES_IO_BYTES, first_symbol: 0 };
        check(unsafe { ffi::gdx_cursors_many(self.device.0 .0, &q, starts.as_mut_ptr(), ends.as_mut_ptr()) });
        starts.into_iter().zip(ends).map(move |(s, e)| Cursor {
            index: self, interval: HalfOpenInterval { start: s as usize, end: e as usize } })
    }

    // src/lib.rs:155-161
    pub fn count_many<Q: AsRef<[u8]>>(&self, queries: impl IntoIterator<Item = Q>) -> impl Iterator<Item = usize> {
        let (bytes, offsets) = pack(queries);
        let nq = offsets.len() - 1;
        let mut counts = vec![0u64; nq];
        let q = ffi::gdx_queries { bytes: bytes.as_ptr(), offsets: offsets.as_ptr(), fixed_len: 0, nq: nq as u64,
                                   encoding: ffi::GDX_QUERIES_IO_BYTES, first_symbol: 0 };
        check(unsafe { ffi::gdx_count_many(self.device.0 .0, &q, counts.as_mut_ptr()) });
        counts.into_iter().map(|c| c as usize)
    }

    // src/lib.rs:179-185: an iterator of iterators over the CSR result
    pub fn locate_many<Q: AsRef<[u8]>>(
        &self, queries: impl IntoIterator<Item = Q>,
    ) -> impl Iterator<Item: Iterator<Item = Hit>> {
        let (bytes, offsets) = pack(queries);
        let nq = offsets.len() - 1;
        let mut hit_offsets = vec![0u64; nq + 1];
        let (mut hits, mut n) = (std::ptr::null_mut(), 0u64);
        let q = ffi::gdx_queries { bytes: bytes.as_ptr(), offsets: offsets.as_ptr(), fixed_len: 0, nq: nq as u64,
                                   encoding: ffi::GDX_QUERIES_IO_BYTES, first_symbol: 0 };
        check(unsafe { ffi::gdx_locate_many(self.device.0 .0, &q, hit_offsets.as_mut_ptr(), &mut hits, &mut n) });
        let owned: Vec<Hit> = unsafe { std::slice::from_raw_parts(hits, n as usize) }
            .iter().map(|h| Hit { text_id: h.text_id as usize, position: h.position as usize }).collect();
        unsafe { ffi::gdx_free_hits(self.device.0 .0, hits) };
        let owned = std::rc::Rc::new(owned);
        (0..nq).map(move |i| {
            let (a, b, o) = (hit_offsets[i] as usize, hit_offsets[i + 1] as usize, owned.clone());
            (a..b).map(move |k| o[k])
        })
    }
}
