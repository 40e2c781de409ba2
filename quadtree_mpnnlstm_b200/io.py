"""Forecast writer of the experiment scripts (ice_exp.py:229-241, ice_inf.py likewise): the reference packs the validation
forecasts into an ``xarray.Dataset`` -- variables ``y_hat`` / ``y_true`` on (launch_date, timestep, latitude, longitude),
coordinates ``longitude``, ``latitude``, ``launch_date`` (datetimes), ``timestep`` = 1..T_out -- and calls ``to_netcdf``.

xarray / netCDF4 are not part of this image, so the file is written directly as NetCDF-3 (classic format, 64-bit offsets)
through ``scipy.io.netcdf_file`` with the same names, dimensions and CF time encoding (``days since 1970-01-01``), i.e.
``xarray.open_dataset`` returns the dataset the reference would have written.  Host-side I/O, outside the hot path; the
forecasts arrive as the single device -> host copy ``NextFramePredictorS2S.predict`` makes.
"""
from __future__ import annotations

import datetime

import numpy as np


def _days_since_epoch(launch_dates):
    out = []
    for d in launch_dates:
        if isinstance(d, (np.datetime64,)):
            out.append(d.astype("datetime64[ns]").astype("int64") / 8.64e13)
        elif isinstance(d, datetime.datetime):
            out.append((d - datetime.datetime(1970, 1, 1)).total_seconds() / 86400.0)
        else:                                   # int64 nanoseconds, as the datasets serve them (ice_dataset.py:60)
            out.append(int(d) / 8.64e13)
    return np.asarray(out, dtype=np.float64)


def write_forecasts(path, y_hat, y_true, latitude, longitude, launch_dates, output_timesteps=None):
    """``y_hat`` / ``y_true``: [n_launch_dates, T_out, H, W] or [..., 1] (numpy or torch); ``latitude`` [H], ``longitude`` [W];
    ``launch_dates``: datetimes, ``numpy.datetime64`` or int64 nanoseconds.  Writes ``path`` (NetCDF-3) and returns it."""
    from scipy.io import netcdf_file

    def arr(a):
        a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
        return a[..., 0] if a.ndim == 5 else a

    y_hat = arr(y_hat).astype(np.float32)
    y_true = arr(y_true).astype(np.float32) if y_true is not None else None
    n, T, H, W = y_hat.shape
    T_out = int(output_timesteps) if output_timesteps is not None else T
    assert T == T_out and len(launch_dates) == n and len(latitude) == H and len(longitude) == W
    with netcdf_file(path, "w", version=2) as f:
        for name, size in (("launch_date", n), ("timestep", T), ("latitude", H), ("longitude", W)):
            f.createDimension(name, size)
        v = f.createVariable("launch_date", "f8", ("launch_date",))
        v[:] = _days_since_epoch(launch_dates)
        v.units = "days since 1970-01-01 00:00:00"
        v.calendar = "proleptic_gregorian"
        v = f.createVariable("timestep", "i4", ("timestep",))
        v[:] = np.arange(1, T + 1, dtype=np.int32)
        v = f.createVariable("latitude", "f8", ("latitude",))
        v[:] = np.asarray(latitude, dtype=np.float64)
        v = f.createVariable("longitude", "f8", ("longitude",))
        v[:] = np.asarray(longitude, dtype=np.float64)
        dims = ("launch_date", "timestep", "latitude", "longitude")
        v = f.createVariable("y_hat", "f4", dims)
        v[:] = y_hat
        if y_true is not None:
            v = f.createVariable("y_true", "f4", dims)
            v[:] = y_true
    return path
