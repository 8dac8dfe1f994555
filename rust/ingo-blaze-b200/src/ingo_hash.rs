//! `PoseidonClient` (reference `src/ingo_hash/{poseidon_api.rs,utils.rs}`).
use crate::driver_client::*;
use crate::error::*;
use crate::ffi;
use std::ffi::CString;

pub fn num_of_elements_oct_tree(tree_height: u32) -> u32 { (0..tree_height).map(|i| 8u32.pow(tree_height - i - 1)).sum() }
pub fn num_of_elements_in_base_layer(tree_height: u32) -> u32 { 8u32.pow(tree_height - 1) }

#[repr(u8)] #[derive(PartialEq, Eq, Copy, Clone)] pub enum TreeMode { TreeC, TreeD }
impl TreeMode { pub fn value(tree_mode: TreeMode) -> u32 { match tree_mode { TreeMode::TreeC => 0, TreeMode::TreeD => 1 } } }

pub enum Hash { Poseidon }
#[derive(Clone)] pub struct PoseidonInitializeParameters { pub tree_height: u32, pub tree_mode: TreeMode, pub instruction_path: String }
pub struct PoseidonResult { pub hash_byte: [u8; 32], pub hash_id: u32, pub layer_id: u32 }

impl PoseidonResult {
    /// 64-byte records: hash[32] || meta[32], meta = LE(hash_id | layer_id << 30).
    pub fn parse_poseidon_hash_results(data: Vec<u8>) -> Vec<PoseidonResult> {
        data.chunks_exact(64).map(|rec| {
            let mut hash_byte = [0u8; 32];
            hash_byte.copy_from_slice(&rec[..32]);
            let meta = u64::from_le_bytes(rec[32..40].try_into().unwrap());
            PoseidonResult { hash_byte, hash_id: (meta & 0x3fff_ffff) as u32, layer_id: ((meta >> 30) & 0x3ff) as u32 }
        }).collect()
    }
}

pub struct PoseidonClient { h: *mut ffi::bz_poseidon, pub dclient: DriverClient }
unsafe impl Send for PoseidonClient {}

impl<'a> DriverPrimitive<Hash, PoseidonInitializeParameters, &'a [u8], Vec<PoseidonResult>> for PoseidonClient {
    fn new(_ptype: Hash, dclient: DriverClient) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::bz_poseidon_new(dclient.h, 0, &mut h) }).unwrap();
        PoseidonClient { h, dclient }
    }
    fn loaded_binary_parameters(&self) -> Vec<u32> {
        let mut v = [0u32; 2];
        check(unsafe { ffi::bz_poseidon_loaded_binary_parameters(self.h, v.as_mut_ptr()) }).unwrap();
        v.to_vec()
    }
    fn initialize(&self, param: PoseidonInitializeParameters) -> Result<()> {
        let path = CString::new(param.instruction_path.clone()).map_err(|_| DriverClientError::InvalidPrimitiveParam)?;
        check(unsafe { ffi::bz_poseidon_initialize(self.h, param.tree_height, TreeMode::value(param.tree_mode) as i32, path.as_ptr()) })
    }
    fn start_process(&self, _param: Option<usize>) -> Result<()> { Ok(()) } // todo!() in the reference
    fn set_data(&self, input: &'a [u8]) -> Result<()> { check(unsafe { ffi::bz_poseidon_set_data(self.h, input.as_ptr(), input.len()) }) }
    fn wait_result(&self) -> Result<()> { Ok(()) } // todo!() in the reference
    fn result(&self, expected_result: Option<usize>) -> Result<Option<Vec<PoseidonResult>>> {
        let expected = expected_result.ok_or(DriverClientError::InvalidPrimitiveParam)?;
        let cap = expected + self.get_num_of_pending_results()? as usize + 8;
        let mut raw = vec![0u8; cap * 64];
        let mut got = 0usize;
        check(unsafe { ffi::bz_poseidon_result(self.h, expected, raw.as_mut_ptr(), cap, &mut got) })?;
        raw.truncate(got * 64);
        Ok(Some(PoseidonResult::parse_poseidon_hash_results(raw)))
    }
}

impl PoseidonClient {
    pub fn get_last_element_sent_to_ring(&self) -> Result<u32> { let mut v = 0; check(unsafe { ffi::bz_poseidon_get_last_element_sent_to_ring(self.h, &mut v) }).map(|_| v) }
    pub fn get_num_of_pending_results(&self) -> Result<u32> { let mut v = 0; check(unsafe { ffi::bz_poseidon_get_num_of_pending_results(self.h, &mut v) }).map(|_| v) }
    pub fn get_raw_results(&self, num_of_results: u32) -> Result<Vec<u8>> {
        let mut res = vec![0u8; 64 * num_of_results as usize];
        check(unsafe { ffi::bz_poseidon_get_raw_results(self.h, num_of_results, res.as_mut_ptr()) })?;
        Ok(res)
    }
    pub fn get_last_hash_sent_to_host(&self) -> Result<u32> { let mut v = 0; check(unsafe { ffi::bz_poseidon_get_last_hash_sent_to_host(self.h, &mut v) }).map(|_| v) }
    pub fn log_api_values(&self) {}
}
impl Drop for PoseidonClient { fn drop(&mut self) { unsafe { ffi::bz_poseidon_free(self.h); } } }
