//! Raw bindings of include/blaze_b200.h (only what the safe wrappers need).
#![allow(non_camel_case_types)]
use std::os::raw::c_char;

#[repr(C)] pub struct bz_dclient { _p: [u8; 0] }
#[repr(C)] pub struct bz_msm { _p: [u8; 0] }
#[repr(C)] pub struct bz_ntt { _p: [u8; 0] }
#[repr(C)] pub struct bz_poseidon { _p: [u8; 0] }

extern "C" {
    pub fn bz_last_error() -> *const c_char;
    pub fn bz_dclient_new(id: *const c_char, card_type: i32, out: *mut *mut bz_dclient) -> i32;
    pub fn bz_dclient_free(dc: *mut bz_dclient) -> i32;
    pub fn bz_dclient_reset(dc: *mut bz_dclient) -> i32;
    pub fn bz_dclient_dma_write(dc: *mut bz_dclient, base: u64, offset: u64, data: *const u8, len: usize) -> i32;
    pub fn bz_dclient_dma_read(dc: *mut bz_dclient, base: u64, offset: u64, out: *mut u8, len: usize) -> i32;
    pub fn bz_dclient_firewalls_status(dc: *mut bz_dclient, blocked: *mut u32) -> i32;
    pub fn bz_dclient_unblock_firewalls(dc: *mut bz_dclient) -> i32;
    pub fn bz_dclient_initialize_cms(dc: *mut bz_dclient) -> i32;
    pub fn bz_dclient_reset_sensor_data(dc: *mut bz_dclient) -> i32;
    pub fn bz_dclient_setup_before_load_binary(dc: *mut bz_dclient) -> i32;
    pub fn bz_dclient_load_binary(dc: *mut bz_dclient, image: *const u8, len: usize) -> i32;
    pub fn bz_dclient_device_count(dc: *mut bz_dclient, n: *mut u32) -> i32;
    pub fn bz_comm_unique_id(out: *mut u8) -> i32;
    pub fn bz_dclient_comm_init(dc: *mut bz_dclient, rank: i32, world: i32, unique_id: *const u8) -> i32;

    pub fn bz_msm_new(dc: *mut bz_dclient, curve: i32, mem_type: i32, is_precompute: i32, out: *mut *mut bz_msm) -> i32;
    pub fn bz_msm_free(m: *mut bz_msm) -> i32;
    pub fn bz_msm_loaded_binary_parameters(m: *mut bz_msm, out: *mut u32) -> i32;
    pub fn bz_msm_initialize(m: *mut bz_msm, nof_elements: u32, has_hbm: i32, hbm_addr: u64, hbm_off: u64) -> i32;
    pub fn bz_msm_start_process(m: *mut bz_msm) -> i32;
    pub fn bz_msm_set_data(m: *mut bz_msm, points: *const u8, points_len: usize, scalars: *const u8, scalars_len: usize,
                           nof_elements: u32, has_hbm: i32, hbm_addr: u64, hbm_off: u64) -> i32;
    pub fn bz_msm_wait_result(m: *mut bz_msm) -> i32;
    pub fn bz_msm_result(m: *mut bz_msm, out: *mut u8, out_len: usize, label: *mut u32) -> i32;
    pub fn bz_msm_task_label(m: *mut bz_msm, label: *mut u32) -> i32;
    pub fn bz_msm_nof_elements(m: *mut bz_msm, n: *mut u32) -> i32;
    pub fn bz_msm_is_msm_engine_ready(m: *mut bz_msm, ready: *mut u32) -> i32;
    pub fn bz_msm_load_data_to_hbm(m: *mut bz_msm, points: *const u8, len: usize, addr: u64, offset: u64) -> i32;
    pub fn bz_msm_get_data_from_hbm(m: *mut bz_msm, out: *mut u8, len: usize, addr: u64, offset: u64) -> i32;
    pub fn bz_msm_sizes(m: *mut bz_msm, scalar: *mut u32, point: *mut u32, result: *mut u32, factor: *mut u32) -> i32;
    pub fn bz_msm_set_precompute(m: *mut bz_msm, mode: i32) -> i32;
    pub fn bz_msm_get_api(m: *mut bz_msm, regs: *mut u32, n_words: usize) -> i32;

    pub fn bz_ntt_new(dc: *mut bz_dclient, ntt_type: i32, out: *mut *mut bz_ntt) -> i32;
    pub fn bz_ntt_free(t: *mut bz_ntt) -> i32;
    pub fn bz_ntt_loaded_binary_parameters(t: *mut bz_ntt, out: *mut u32) -> i32;
    pub fn bz_ntt_initialize(t: *mut bz_ntt) -> i32;
    pub fn bz_ntt_set_data(t: *mut bz_ntt, buf_host: usize, data: *const u8, len: usize) -> i32;
    pub fn bz_ntt_start_process(t: *mut bz_ntt, buf_kernel: usize) -> i32;
    pub fn bz_ntt_wait_result(t: *mut bz_ntt) -> i32;
    pub fn bz_ntt_result(t: *mut bz_ntt, buf_num: usize, out: *mut u8, out_len: usize) -> i32;

    pub fn bz_poseidon_new(dc: *mut bz_dclient, hash_type: i32, out: *mut *mut bz_poseidon) -> i32;
    pub fn bz_poseidon_free(p: *mut bz_poseidon) -> i32;
    pub fn bz_poseidon_loaded_binary_parameters(p: *mut bz_poseidon, out: *mut u32) -> i32;
    pub fn bz_poseidon_initialize(p: *mut bz_poseidon, tree_height: u32, tree_mode: i32, path: *const c_char) -> i32;
    pub fn bz_poseidon_set_data(p: *mut bz_poseidon, input: *const u8, len: usize) -> i32;
    pub fn bz_poseidon_result(p: *mut bz_poseidon, expected: usize, out: *mut u8, cap_records: usize, n_out: *mut usize) -> i32;
    pub fn bz_poseidon_get_last_element_sent_to_ring(p: *mut bz_poseidon, id: *mut u32) -> i32;
    pub fn bz_poseidon_get_num_of_pending_results(p: *mut bz_poseidon, n: *mut u32) -> i32;
    pub fn bz_poseidon_get_raw_results(p: *mut bz_poseidon, n: u32, out: *mut u8) -> i32;
    pub fn bz_poseidon_get_last_hash_sent_to_host(p: *mut bz_poseidon, id: *mut u32) -> i32;
}
