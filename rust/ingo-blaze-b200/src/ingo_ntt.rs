//! `NTTClient` (reference `src/ingo_ntt/ntt_api.rs`): 2^27 elements of 32 bytes, two buffer slots.
use crate::driver_client::*;
use crate::error::*;
use crate::ffi;

pub enum NTT { Ntt }
pub struct NttInit {}
#[derive(Debug, Clone)] pub struct NTTInput { pub buf_host: usize, pub data: Vec<u8> }

pub const NTT_BYTES: usize = (1usize << 27) * 32;

pub struct NTTClient { h: *mut ffi::bz_ntt, pub driver_client: DriverClient }
unsafe impl Send for NTTClient {}
unsafe impl Sync for NTTClient {}

impl DriverPrimitive<NTT, NttInit, NTTInput, Vec<u8>> for NTTClient {
    fn new(_ptype: NTT, dclient: DriverClient) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::bz_ntt_new(dclient.h, 0, &mut h) }).unwrap();
        NTTClient { h, driver_client: dclient }
    }
    fn loaded_binary_parameters(&self) -> Vec<u32> {
        let mut v = [0u32; 2];
        check(unsafe { ffi::bz_ntt_loaded_binary_parameters(self.h, v.as_mut_ptr()) }).unwrap();
        v.to_vec()
    }
    fn initialize(&self, _: NttInit) -> Result<()> { check(unsafe { ffi::bz_ntt_initialize(self.h) }) }
    fn start_process(&self, buf_kernel: Option<usize>) -> Result<()> {
        check(unsafe { ffi::bz_ntt_start_process(self.h, buf_kernel.ok_or(DriverClientError::InvalidPrimitiveParam)?) })
    }
    fn set_data(&self, input: NTTInput) -> Result<()> {
        check(unsafe { ffi::bz_ntt_set_data(self.h, input.buf_host, input.data.as_ptr(), input.data.len()) })
    }
    fn wait_result(&self) -> Result<()> { check(unsafe { ffi::bz_ntt_wait_result(self.h) }) }
    fn result(&self, buf_num: Option<usize>) -> Result<Option<Vec<u8>>> {
        let mut res = vec![0u8; NTT_BYTES];
        check(unsafe { ffi::bz_ntt_result(self.h, buf_num.ok_or(DriverClientError::InvalidPrimitiveParam)?, res.as_mut_ptr(), res.len()) })?;
        Ok(Some(res))
    }
}
impl Drop for NTTClient { fn drop(&mut self) { unsafe { ffi::bz_ntt_free(self.h); } } }
