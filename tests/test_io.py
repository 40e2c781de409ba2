"""the netCDF forecast writer (ice_exp.py:229-241) -- names, dimensions, coordinates and CF time encoding of
the file the reference writes through xarray, read back here with scipy's NetCDF-3 reader."""
import datetime

import numpy as np
import torch


def test_forecast_file_round_trip(tmp_path):
    from scipy.io import netcdf_file
    from quadtree_mpnnlstm_b200.io import write_forecasts
    rng = np.random.default_rng(0)
    n, T, H, W = 3, 4, 5, 6
    y_hat = rng.random((n, T, H, W, 1)).astype(np.float32)
    y_hat[0, 0, 0, 0, 0] = np.nan                                # masked pixels of the pixel-wise mesh
    y_true = torch.from_numpy(rng.random((n, T, H, W, 1)).astype(np.float32))
    t0 = np.datetime64("2016-02-27").astype("datetime64[ns]").astype("int64")
    dates = [t0 + k * 86_400_000_000_000 for k in range(n)]
    path = write_forecasts(str(tmp_path / "valpredictions.nc"), y_hat, y_true, np.linspace(50, 70, H), np.linspace(-95, -65, W),
                           dates, output_timesteps=T)
    with netcdf_file(path, "r", mmap=False) as f:
        assert set(f.dimensions) == {"launch_date", "timestep", "latitude", "longitude"}
        assert f.variables["y_hat"].dimensions == ("launch_date", "timestep", "latitude", "longitude")
        got = f.variables["y_hat"][:]
        assert np.array_equal(np.isnan(got), np.isnan(y_hat[..., 0])) and np.allclose(np.nan_to_num(got), np.nan_to_num(y_hat[..., 0]))
        assert np.allclose(f.variables["y_true"][:], y_true[..., 0].numpy())
        assert list(f.variables["timestep"][:]) == [1, 2, 3, 4]
        days = f.variables["launch_date"][:]
        assert f.variables["launch_date"].units.decode().startswith("days since 1970-01-01")
        back = [datetime.datetime(1970, 1, 1) + datetime.timedelta(days=float(d)) for d in days]
        assert back == [datetime.datetime(2016, 2, 27), datetime.datetime(2016, 2, 28), datetime.datetime(2016, 2, 29)]
        assert np.allclose(f.variables["latitude"][:], np.linspace(50, 70, H))
