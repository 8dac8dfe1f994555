//! Same error type and variants as the reference (`src/error.rs:4-32`), filled from `bz_status`.
use std::io;
use thiserror::Error;

pub type Result<T> = std::result::Result<T, DriverClientError>;

#[derive(Error, Debug)]
pub enum DriverClientError {
    #[error("failed to write data in offset {:?}", offset)]
    WriteError { offset: String, #[source] source: io::Error },
    #[error("failed to read data from offset {:?}", offset)]
    ReadError { offset: String, #[source] source: io::Error },
    #[error("hbicap doesn't ready to work")]
    HBICAPNotReady,
    #[error("failed to get driver primitive param")]
    InvalidPrimitiveParam,
    #[error("failed to load instruction set from: {:?}", path)]
    LoadFailed { path: String },
    #[error("failed open file")]
    FileError(#[from] io::Error),
    #[error("unknown driver client error")]
    Unknown,
}

fn last_message() -> String {
    unsafe { std::ffi::CStr::from_ptr(crate::ffi::bz_last_error()) }.to_string_lossy().into_owned()
}

/// Map a `bz_status` (include/blaze_b200.h) onto the reference's variants.
pub fn check(rc: i32) -> Result<()> {
    let io = || io::Error::new(io::ErrorKind::Other, last_message());
    match rc {
        0 => Ok(()),
        -1 => Err(DriverClientError::WriteError { offset: "b200".into(), source: io() }),
        -2 | -10 => Err(DriverClientError::ReadError { offset: "b200".into(), source: io() }),
        -3 => Err(DriverClientError::HBICAPNotReady),
        -4 => Err(DriverClientError::InvalidPrimitiveParam),
        -6 => Err(DriverClientError::LoadFailed { path: last_message() }),
        -7 | -9 => Err(DriverClientError::FileError(io())),
        _ => Err(DriverClientError::Unknown),
    }
}
