//! `MSMClient` and its parameter types (reference `src/ingo_msm/{msm_api.rs,msm_cfg.rs}`).
use crate::driver_client::*;
use crate::error::*;
use crate::ffi;

#[derive(Debug, PartialEq)] pub enum Curve { BLS377, BLS381, BN254 }
#[derive(Debug, PartialEq, Eq, Clone, Copy)] pub enum PointMemoryType { HBM, DMA }

pub struct MSMInit { pub mem_type: PointMemoryType, pub is_precompute: bool, pub curve: Curve }
#[derive(Debug, Copy, Clone)] pub struct MSMParams { pub nof_elements: u32, pub hbm_point_addr: Option<(u64, u64)> }
pub struct MSMInput { pub points: Option<Vec<u8>>, pub scalars: Vec<u8>, pub params: MSMParams }
#[derive(Debug, Clone)] pub struct MSMResult { pub result: Vec<u8>, pub result_label: u32 }

pub const PRECOMPUTE_FACTOR_BASE: u32 = 1;
pub const PRECOMPUTE_FACTOR: u32 = 8;

pub struct MSMClient { h: *mut ffi::bz_msm, result_point_size: usize, pub driver_client: DriverClient }
unsafe impl Send for MSMClient {}
unsafe impl Sync for MSMClient {}

fn hbm(p: &MSMParams) -> (i32, u64, u64) { p.hbm_point_addr.map_or((0, 0, 0), |(a, o)| (1, a, o)) }

impl DriverPrimitive<MSMInit, MSMParams, MSMInput, MSMResult> for MSMClient {
    fn new(init: MSMInit, dclient: DriverClient) -> Self {
        let curve = match init.curve { Curve::BLS377 => 0, Curve::BN254 => 1, Curve::BLS381 => 2 };
        let mem = match init.mem_type { PointMemoryType::HBM => 0, PointMemoryType::DMA => 1 };
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::bz_msm_new(dclient.h, curve, mem, init.is_precompute as i32, &mut h) }).unwrap();
        let (mut s, mut p, mut r, mut f) = (0u32, 0u32, 0u32, 0u32);
        check(unsafe { ffi::bz_msm_sizes(h, &mut s, &mut p, &mut r, &mut f) }).unwrap();
        MSMClient { h, result_point_size: r as usize, driver_client: dclient }
    }
    fn loaded_binary_parameters(&self) -> Vec<u32> {
        let mut v = [0u32; 2];
        check(unsafe { ffi::bz_msm_loaded_binary_parameters(self.h, v.as_mut_ptr()) }).unwrap();
        v.to_vec()
    }
    fn initialize(&self, params: MSMParams) -> Result<()> {
        let (has, a, o) = hbm(&params);
        check(unsafe { ffi::bz_msm_initialize(self.h, params.nof_elements, has, a, o) })
    }
    fn start_process(&self, _param: Option<usize>) -> Result<()> { check(unsafe { ffi::bz_msm_start_process(self.h) }) }
    fn set_data(&self, data: MSMInput) -> Result<()> {
        let (pp, pl) = data.points.as_ref().map_or((std::ptr::null(), 0), |p| (p.as_ptr(), p.len()));
        let (has, a, o) = hbm(&data.params);
        check(unsafe { ffi::bz_msm_set_data(self.h, pp, pl, data.scalars.as_ptr(), data.scalars.len(), data.params.nof_elements, has, a, o) })
        // `data` drops here: the library has finished reading the buffers (move-in semantics kept)
    }
    fn wait_result(&self) -> Result<()> { check(unsafe { ffi::bz_msm_wait_result(self.h) }) }
    fn result(&self, _param: Option<usize>) -> Result<Option<MSMResult>> {
        let mut out = vec![0u8; self.result_point_size];
        let mut label = 0u32;
        check(unsafe { ffi::bz_msm_result(self.h, out.as_mut_ptr(), out.len(), &mut label) })?;
        Ok(Some(MSMResult { result: out, result_label: label }))
    }
}

impl MSMClient {
    pub fn task_label(&self) -> Result<u32> { let mut v = 0; check(unsafe { ffi::bz_msm_task_label(self.h, &mut v) }).map(|_| v) }
    pub fn nof_elements(&self) -> Result<u32> { let mut v = 0; check(unsafe { ffi::bz_msm_nof_elements(self.h, &mut v) }).map(|_| v) }
    pub fn is_msm_engine_ready(&self) -> Result<u32> { let mut v = 0; check(unsafe { ffi::bz_msm_is_msm_engine_ready(self.h, &mut v) }).map(|_| v) }
    pub fn load_data_to_hbm(&self, points: &[u8], addr: u64, offset: u64) -> Result<()> {
        check(unsafe { ffi::bz_msm_load_data_to_hbm(self.h, points.as_ptr(), points.len(), addr, offset) })
    }
    pub fn get_data_from_hbm(&self, data_len: usize, addr: u64, offset: u64) -> Result<Vec<u8>> {
        let mut res = vec![0u8; data_len];
        check(unsafe { ffi::bz_msm_get_data_from_hbm(self.h, res.as_mut_ptr(), data_len, addr, offset) })?;
        Ok(res)
    }
    /// B200 addition: 0 = never, 1 = on reuse (default), 2 = always derive the table of window multiples
    /// (2^(c w) P) from an HBM-resident point set so that all windows share one bucket set.
    pub fn set_precompute(&self, mode: i32) -> Result<()> { check(unsafe { ffi::bz_msm_set_precompute(self.h, mode) }) }
    /// `msm_api.rs:324-330`: reads every `INGO_MSM_ADDR` register (word index = offset / 4, `msm_hw_code.rs:6-55`);
    /// the reference discards the values, so does this -- `get_api_values` returns them.
    pub fn get_api(&self) { let _ = self.get_api_values(); }
    pub fn get_api_values(&self) -> Result<Vec<u32>> {
        let mut regs = vec![0u32; 82];
        check(unsafe { ffi::bz_msm_get_api(self.h, regs.as_mut_ptr(), regs.len()) })?;
        Ok(regs)
    }
}
impl Drop for MSMClient { fn drop(&mut self) { unsafe { ffi::bz_msm_free(self.h); } } }
