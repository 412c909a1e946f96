// NOT COMPILED in this repository's environment (no rustc/cargo in the image).
// Reference-side binding of libgenedex_b200.so, kept in sync with INTEGRATION.md and include/genedex_b200.h.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub type gdx_status = i32;            // 0 = GDX_OK, 1 = GDX_ERR_INVALID_SYMBOL, ...
#[repr(C)] pub struct gdx_index { _private: [u8; 0] }

#[repr(C)] pub struct gdx_alphabet {
    pub io_to_dense: [u8; 256],
    pub num_dense_symbols: u32,
    pub num_searchable_dense_symbols: u32,
}
#[repr(C)] pub struct gdx_config {
    pub storage: u32,                      // 0 = i32, 1 = u32, 2 = i64
    pub suffix_array_sampling_rate: u32,
    pub lookup_table_depth: u32,
    pub performance_priority: u32,
    pub construction: u32,                 // 0 = host SA-IS, 1 = GPU prefix doubling, 2 = auto
    pub device: i32,
    pub flags: u32,
}
#[repr(C)] #[derive(Clone, Copy)] pub struct gdx_hit { pub text_id: u64, pub position: u64 }
#[repr(C)] pub struct gdx_queries {
    pub bytes: *const u8,
    pub offsets: *const u64,               // nq + 1 entries, or null => fixed_len
    pub fixed_len: u64,
    pub nq: u64,
}
#[repr(C)] pub struct gdx_parts {           // an index the crate built itself, in its own layout
    pub alphabet: gdx_alphabet,
    pub storage: u32,
    pub text_len: u64,
    pub count: *const u64,
    pub interleaved_blocks: *const u64,
    pub interleaved_block_offsets: *const u16,
    pub sampled_suffix_array: *const u64,
    pub sampling_rate: u32,
    pub text_border_rows: *const u64,
    pub text_border_positions: *const u64,
    pub num_text_borders: u64,
    pub sentinel_indices: *const u64,
    pub num_texts: u64,
    pub lookup_table_depth: u32,
}

extern "C" {
    pub fn gdx_last_error_message() -> *const c_char;
    pub fn gdx_last_error_query() -> u64;
    pub fn gdx_index_build(texts: *const u8, text_offsets: *const u64, num_texts: u64,
                           alphabet: *const gdx_alphabet, config: *const gdx_config,
                           out: *mut *mut gdx_index) -> gdx_status;
    pub fn gdx_index_create_from_parts(parts: *const gdx_parts, device: i32,
                                       out: *mut *mut gdx_index) -> gdx_status;
    pub fn gdx_index_destroy(idx: *mut gdx_index);
    pub fn gdx_cursors_many(idx: *const gdx_index, q: *const gdx_queries,
                            starts: *mut u64, ends: *mut u64) -> gdx_status;
    pub fn gdx_count_many(idx: *const gdx_index, q: *const gdx_queries, counts: *mut u64) -> gdx_status;
    pub fn gdx_locate_many(idx: *const gdx_index, q: *const gdx_queries, hit_offsets: *mut u64,
                           hits: *mut *mut gdx_hit, num_hits: *mut u64) -> gdx_status;
    pub fn gdx_locate_intervals(idx: *const gdx_index, starts: *const u64, ends: *const u64, n: u64,
                                hit_offsets: *mut u64, hits: *mut *mut gdx_hit,
                                num_hits: *mut u64) -> gdx_status;
    pub fn gdx_free_hits(idx: *const gdx_index, hits: *mut gdx_hit);
    pub fn gdx_extend_many(idx: *const gdx_index, starts: *mut u64, ends: *mut u64,
                           io_symbols: *const u8, n: u64) -> gdx_status;
    pub fn gdx_host_alloc(bytes: u64, out: *mut *mut c_void) -> gdx_status;
    pub fn gdx_host_free(p: *mut c_void);
}
