"""Result / intermediate IO.  The reference writes `<sample>.hdf` with h5py
(vote_instances.py:542-554) and caches blocks in zarr; both packages are
optional here, `.npz` is the always-available format."""
import os

import numpy as np


def write_result(path_noext, datasets, output_format='hdf'):
    if output_format == 'hdf':
        try:
            import h5py
        except ImportError:
            output_format = 'npz'
        else:
            with h5py.File(path_noext + '.hdf', 'w') as f:
                for k, v in datasets.items():
                    f.create_dataset(k, data=v, compression='gzip')
                    f[k].attrs['offset'] = (0, 0, 0)
                    f[k].attrs['resolution'] = (1, 1, 1)
            return path_noext + '.hdf'
    np.savez_compressed(path_noext + '.npz', **datasets)
    return path_noext + '.npz'


def open_zarr(path, mode='r'):
    try:
        import zarr
    except ImportError as e:
        raise RuntimeError("reading %s needs the zarr package; pass arrays to "
                           "do_block / stitch_patch_graph.stitch_arrays instead"
                           % path) from e
    return zarr.open(path, mode)
