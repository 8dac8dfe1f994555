//! `DriverClient`, `DriverConfig`, `CardType`, `DriverPrimitive` (reference `src/driver_client/*`).
use crate::error::*;
use crate::ffi;
use std::ffi::CString;

pub enum CardType { C1100, B200 }

#[derive(Copy, Clone, Debug)]
pub struct DriverConfig { pub(crate) card_type: i32 }
impl DriverConfig {
    pub fn driver_client_cfg(card_type: CardType) -> Self {
        DriverConfig { card_type: match card_type { CardType::C1100 => 0, CardType::B200 => 1 } }
    }
}

#[allow(non_camel_case_types)]
pub enum DMA_RW { OFFSET = 0 }

/// The 7-method trait of the reference (`dclient.rs:28-46`).
pub trait DriverPrimitive<T, P, I, O> {
    fn new(ptype: T, dclient: DriverClient) -> Self;
    fn loaded_binary_parameters(&self) -> Vec<u32>;
    fn initialize(&self, param: P) -> Result<()>;
    fn set_data(&self, input: I) -> Result<()>;
    fn start_process(&self, param: Option<usize>) -> Result<()>;
    fn wait_result(&self) -> Result<()>;
    fn result(&self, param: Option<usize>) -> Result<Option<O>>;
}

pub struct DriverClient { pub(crate) h: *mut ffi::bz_dclient, pub cfg: DriverConfig }
unsafe impl Send for DriverClient {}
unsafe impl Sync for DriverClient {}

impl DriverClient {
    /// `id` = the slot string of the reference = CUDA device ordinal (`dclient.rs:79-86`); a comma-separated list
    /// ("0,1,2,3,4,5,6,7") opens ONE client over several GPUs: an `MSMClient` built on it shards bases and scalars over
    /// them and still returns one `MSMResult` per task.
    pub fn new(id: &str, cfg: DriverConfig) -> Self {
        let mut h = std::ptr::null_mut();
        let cid = CString::new(id).unwrap();
        check(unsafe { ffi::bz_dclient_new(cid.as_ptr(), cfg.card_type, &mut h) }).unwrap(); // the reference unwrap()s the open too
        DriverClient { h, cfg }
    }
    /// Number of GPUs behind this client.
    pub fn device_count(&self) -> Result<u32> { let mut n = 0; check(unsafe { ffi::bz_dclient_device_count(self.h, &mut n) }).map(|_| n) }
    /// One process per GPU: 128-byte NCCL id, made on one rank and handed to all of them.
    pub fn comm_unique_id() -> Result<[u8; 128]> { let mut id = [0u8; 128]; check(unsafe { ffi::bz_comm_unique_id(id.as_mut_ptr()) }).map(|_| id) }
    /// One process per GPU: this client becomes rank `rank` of `world`; MSM results are summed over the ranks on the device.
    pub fn comm_init(&self, rank: i32, world: i32, unique_id: &[u8; 128]) -> Result<()> {
        check(unsafe { ffi::bz_dclient_comm_init(self.h, rank, world, unique_id.as_ptr()) })
    }
    pub fn reset(&self) -> Result<()> { check(unsafe { ffi::bz_dclient_reset(self.h) }) }
    pub fn dma_write(&self, base_address: u64, offset: u64, data: &[u8]) -> Result<()> {
        check(unsafe { ffi::bz_dclient_dma_write(self.h, base_address, offset, data.as_ptr(), data.len()) })
    }
    pub fn dma_read(&self, base_address: u64, offset: u64, data: &mut [u8]) -> Result<()> {
        check(unsafe { ffi::bz_dclient_dma_read(self.h, base_address, offset, data.as_mut_ptr(), data.len()) })
    }
    pub fn firewalls_status(&self) { let mut m = 0u32; let _ = unsafe { ffi::bz_dclient_firewalls_status(self.h, &mut m) }; }
    pub fn unblock_firewalls(&self) -> Result<()> { check(unsafe { ffi::bz_dclient_unblock_firewalls(self.h) }) }
    pub fn initialize_cms(&self) -> Result<()> { check(unsafe { ffi::bz_dclient_initialize_cms(self.h) }) }
    pub fn reset_sensor_data(&self) -> Result<()> { check(unsafe { ffi::bz_dclient_reset_sensor_data(self.h) }) }
    pub fn setup_before_load_binary(&self) -> Result<()> { check(unsafe { ffi::bz_dclient_setup_before_load_binary(self.h) }) }
    pub fn load_binary(&self, binary: &[u8]) -> Result<u32> {
        check(unsafe { ffi::bz_dclient_load_binary(self.h, binary.as_ptr(), binary.len()) }).map(|_| 0)
    }
}
impl Drop for DriverClient { fn drop(&mut self) { unsafe { ffi::bz_dclient_free(self.h); } } }
