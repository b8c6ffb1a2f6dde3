"""Input side of the assembly stage, mirroring the loader functions of
PatchPerPix/vote_instances/utilVoteInstances.py (loadAffinities :136-251,
getFgThreshold :254-258, maybeLoadNuminst :260-272, loadFg :275-303,
returnFg :306-322, getResKey :325-337).

h5py / zarr are optional in this image: `.npy` / `.npz` always work, `.hdf` /
`.zarr` need the corresponding package and raise a clear error otherwise.
"""
import logging

import numpy as np

logger = logging.getLogger(__name__)


def _expit(x):
    return 1.0 / (1.0 + np.exp(-x.astype(np.float32)))


def getFgThreshold(**kwargs):
    if kwargs.get('fg_thresh_vi', -1) > 0:
        return kwargs['fg_thresh_vi']
    return kwargs['patch_threshold']


def numinst_from_prob(numinst_prob, **kwargs):
    """utilVoteInstances.py:263-271."""
    numinst_prob = np.squeeze(numinst_prob)
    if len(numinst_prob.shape) == 3:
        numinst_prob = np.expand_dims(numinst_prob, axis=1)
    numinst = np.argmax(numinst_prob, axis=0).astype(np.uint8)
    if kwargs.get('numinst_threshs'):
        numinst = np.zeros(numinst_prob.shape[1:], dtype=np.uint8)
        for i in range(len(kwargs['numinst_threshs'])):
            numinst[numinst_prob[i + 1] > kwargs['numinst_threshs'][i]] = i + 1
    return numinst


def maybeLoadNuminst(f, **kwargs):
    if kwargs.get('numinst_key') is not None:
        return numinst_from_prob(np.array(f[kwargs['numinst_key']]), **kwargs)
    return None


def loadFg(f, **kwargs):
    """utilVoteInstances.py:275-303: (foreground bool, key)."""
    aff_key = kwargs['aff_key']
    fg_key = kwargs.get('fg_key', None)
    numinst_key = kwargs.get('numinst_key', None)
    fg_thresh = getFgThreshold(**kwargs)
    if fg_key is not None:
        foreground = np.array(f[fg_key])
        key = fg_key
    elif numinst_key is not None:
        numinst_prob = np.array(f[numinst_key])
        numinst = np.argmax(numinst_prob, axis=0).astype(np.uint8)
        if kwargs.get('numinst_threshs'):
            numinst = np.zeros(numinst_prob.shape[1:], dtype=np.uint8)
            for i in range(len(kwargs['numinst_threshs'])):
                numinst[numinst_prob[i + 1] > kwargs['numinst_threshs'][i]] = i + 1
        foreground = np.expand_dims((numinst > 0).astype(np.float32), axis=0)
        key = numinst_key
    else:
        mid = np.prod(kwargs['patchshape']) // 2
        foreground = np.expand_dims(np.array(f[aff_key][mid]), axis=0)
        key = aff_key
    return foreground > fg_thresh, key


def returnFg(affs, numinst, fg, **kwargs):
    """utilVoteInstances.py:306-322."""
    fg_key = kwargs.get('fg_key', None)
    numinst_key = kwargs.get('numinst_key', None)
    fg_thresh = getFgThreshold(**kwargs)
    if fg_key is not None:
        foreground = np.squeeze(fg)
    elif numinst_key is not None:
        foreground = numinst > 0
    else:
        mid = np.prod(kwargs['patchshape']) // 2
        foreground = affs[mid]
    return foreground > fg_thresh


def getResKey(**kwargs):
    res_ext = '_' + str(kwargs['patch_threshold']).replace('.', '')
    if not kwargs.get('skipThinCover', False):
        res_ext += "_tfgc"
    if kwargs['mws']:
        res_ext += "_mws"
    if kwargs['sample'] < 1.0:
        res_ext += "_smp" + str(kwargs['sample']).replace('.', '')
    return res_ext


def _open(aff_file):
    if aff_file.endswith(".hdf"):
        try:
            import h5py
        except ImportError as e:
            raise RuntimeError("reading %s needs h5py" % aff_file) from e
        return h5py.File(aff_file, 'r')
    if aff_file.endswith(".zarr"):
        from .io_util import open_zarr
        return open_zarr(aff_file)
    raise RuntimeError("unsupported container " + aff_file)


def loadAffinities(aff_file, res_ext, patchshape=None, **kwargs):
    """utilVoteInstances.py:136-251 -> (affinities [P,Z,Y,X], numinst, foreground)."""
    numinst = None
    if aff_file.endswith((".hdf", ".zarr")):
        f = _open(aff_file)
        if 'vote_instances' + res_ext in f.keys():
            logger.info("%s vote_instances %s already computed", aff_file, res_ext)
            return None
        if 'volumes' in f.keys():
            aff_key = kwargs.get('aff_key') or 'volumes/pred_affs'
            kwargs['aff_key'] = aff_key
            arr = f[aff_key]
            rotate = False
            if patchshape is not None:
                P = int(np.prod(patchshape))
                rotate = arr.shape[-1] == P and arr.shape[0] != P
            cz = slice(kwargs.get('crop_z_s', 0), kwargs.get('crop_z_e', None))
            cy = slice(kwargs.get('crop_y_s', 0), kwargs.get('crop_y_e', None))
            cx = slice(kwargs.get('crop_x_s', 0), kwargs.get('crop_x_e', None))
            if len(arr.shape) == 3:
                if rotate:
                    a = np.squeeze(np.array(arr[cy, cx, :]))
                    a = np.ascontiguousarray(np.moveaxis(a, -1, 0))
                else:
                    a = np.squeeze(np.array(arr[:, cy, cx]))
                affinities = np.expand_dims(a, axis=1)
            elif len(arr.shape) == 4:
                if rotate:
                    a = np.squeeze(np.array(arr[cz, cy, cx, :]))
                    affinities = np.ascontiguousarray(np.moveaxis(a, -1, 0))
                else:
                    affinities = np.squeeze(np.array(arr[:, cz, cy, cx]))
            else:
                raise RuntimeError("check dimensions of array %s %s" % (aff_file, aff_key))
        else:
            affinities = np.array(f['images/pred_affs'])
            if affinities.shape[1] != 1:
                affinities = np.expand_dims(affinities, axis=1)
        numinst = maybeLoadNuminst(f, **kwargs)
        foreground, _ = loadFg(f, **dict(kwargs, patchshape=patchshape))
        if hasattr(f, 'close'):
            f.close()
    elif aff_file.endswith("npy"):
        affinities = np.load(aff_file)
        if affinities.shape[1] != 1:
            affinities = np.expand_dims(affinities, axis=1)
        mid = np.prod(patchshape) // 2
        foreground = np.array(affinities[mid]) > getFgThreshold(**kwargs)
        numinst = 1 * foreground
    else:
        raise RuntimeError("invalid affinities file, zarr, hdf or npy")
    if np.min(affinities) < 0 and np.max(affinities) > 1:
        affinities = _expit(affinities)
    return affinities, numinst, foreground
