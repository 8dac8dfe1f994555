//! B200 backend behind blaze's client surface.  Module paths follow the reference (`src/lib.rs:8-13`).
pub mod driver_client;
pub mod error;
pub mod ffi;
pub mod ingo_hash;
pub mod ingo_msm;
pub mod ingo_ntt;
