//! Raw bindings of `include/gpumqtt.h`.  UNTESTED SOURCE (no cargo in the image that built libgpumqtt.so).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct gm_engine {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct gm_config {
    pub struct_size: u32,
    pub device: i32,
    pub max_levels: u32,
    pub flags: u32,
    pub filters_hint: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct gm_span {
    pub off: u32,
    pub cnt: u32,
}

pub const GM_OK: i32 = 0;
pub const GM_ERR_INVALID_ARG: i32 = -1;
pub const GM_ERR_INVALID_TOPIC: i32 = -2;
pub const GM_ERR_CAPACITY: i32 = -3;
pub const GM_ERR_CUDA: i32 = -4;
pub const GM_ERR_TOO_DEEP: i32 = -5;
pub const GM_ERR_NO_DEVICE: i32 = -6;
pub const GM_ERR_TOO_LARGE: i32 = -7;
pub const GM_FLAG_MANUAL_FLUSH: u32 = 1;

#[link(name = "gpumqtt")]
extern "C" {
    pub fn gm_create(cfg: *const gm_config, out: *mut *mut gm_engine) -> i32;
    pub fn gm_destroy(e: *mut gm_engine);
    pub fn gm_last_error(e: *mut gm_engine) -> *const c_char;
    pub fn gm_version() -> *const c_char;
    pub fn gm_sub_add(e: *mut gm_engine, filter: *const u8, len: u32, value: u32, changed: *mut i32) -> i32;
    pub fn gm_sub_remove(e: *mut gm_engine, filter: *const u8, len: u32, value: u32, changed: *mut i32) -> i32;
    pub fn gm_bulk_load(e: *mut gm_engine, blob: *const u8, offsets: *const u32, values: *const u32, n: u64, n_changed: *mut u64) -> i32;
    pub fn gm_flush(e: *mut gm_engine) -> i32;
    pub fn gm_compact(e: *mut gm_engine) -> i32;
    pub fn gm_match_batch(
        e: *mut gm_engine, topics_blob: *const u8, topic_offsets: *const u32, n: u64, out_spans: *mut gm_span,
        out_ids: *mut u32, cap_ids: u64, needed: *mut u64, status: *mut i32,
    ) -> i32;
    pub fn gm_retain_set(e: *mut gm_engine, topic: *const u8, len: u32, value: u32, had_old: *mut i32, old_value: *mut u32) -> i32;
    pub fn gm_retain_remove(e: *mut gm_engine, topic: *const u8, len: u32, had_old: *mut i32, old_value: *mut u32) -> i32;
    pub fn gm_retain_bulk_load(e: *mut gm_engine, blob: *const u8, offsets: *const u32, values: *const u32, n: u64, n_set: *mut u64) -> i32;
    pub fn gm_retain_match_batch(
        e: *mut gm_engine, filters_blob: *const u8, filter_offsets: *const u32, n: u64, out_spans: *mut gm_span,
        out_ids: *mut u32, cap_ids: u64, needed: *mut u64, status: *mut i32,
    ) -> i32;
    pub fn gm_shard_of(topic_or_filter: *const u8, len: u32, n_shards: u32) -> u32;
    pub fn gm_shard_of_batch(blob: *const u8, offsets: *const u32, n: u64, n_shards: u32, out_shard: *mut u32) -> i32;
    pub fn gm_host_alloc(bytes: u64) -> *mut c_void;
    pub fn gm_host_free(p: *mut c_void);
}

// ---- round-2 additions of include/gpumqtt.h (same caveat: untested source) -------------------------------------------
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gm_desc {
    pub r#ref: u32, // cnt == 1: the value itself; 2..65534: values[ref .. ref+cnt); 65535: ranges[ref]
    pub cnt: u32,
}
#[repr(C)]
pub struct gm_values {
    pub values: *const u32,
    pub n_values: u64,
    pub ranges: *const gm_span,
    pub n_ranges: u64,
    pub epoch: u64,
}
#[repr(C)]
pub struct gm_batcher {
    _private: [u8; 0],
}
pub type gm_match_cb = extern "C" fn(user: *mut c_void, cookie: u64, status: i32, ids: *const u32, n_ids: u32);
#[repr(C)]
pub struct gm_batcher_config {
    pub struct_size: u32,
    pub max_batch: u32,
    pub max_wait_us: u32,
    pub dispatchers: u32,
    pub on_match: gm_match_cb,
    pub user: *mut c_void,
}
extern "C" {
    pub fn gm_match_batch_desc(e: *mut gm_engine, topics_blob: *const u8, topic_offsets: *const u32, n: u64, out_spans: *mut gm_span,
                               out_descs: *mut gm_desc, cap_descs: u64, needed: *mut u64, status: *mut i32) -> i32;
    pub fn gm_values_view(e: *mut gm_engine, out: *mut gm_values) -> i32;
    pub fn gm_batcher_create(e: *mut gm_engine, cfg: *const gm_batcher_config, out: *mut *mut gm_batcher) -> i32;
    pub fn gm_submit(b: *mut gm_batcher, topic: *const u8, len: u32, cookie: u64) -> i32;
    pub fn gm_submit_publish(b: *mut gm_batcher, packet: *const u8, len: u32, cookie: u64) -> i32;
    pub fn gm_batcher_drain(b: *mut gm_batcher) -> i32;
    pub fn gm_batcher_destroy(b: *mut gm_batcher);
    pub fn gm_sub_add_tree(e: *mut gm_engine, tree: u32, filter: *const u8, len: u32, value: u32, changed: *mut i32) -> i32;
    pub fn gm_sub_remove_tree(e: *mut gm_engine, tree: u32, filter: *const u8, len: u32, value: u32, changed: *mut i32) -> i32;
    pub fn gm_match_batch_trees(e: *mut gm_engine, topics_blob: *const u8, topic_offsets: *const u32, trees: *const u32, n: u64,
                                out_spans: *mut gm_span, out_ids: *mut u32, cap_ids: u64, needed: *mut u64, status: *mut i32) -> i32;
    pub fn gm_retain_remove_batch(e: *mut gm_engine, blob: *const u8, offsets: *const u32, n: u64, old_values: *mut u32, n_removed: *mut u64) -> i32;
}
