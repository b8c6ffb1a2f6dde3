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
    """instance-count class per voxel from the class probabilities [C,(Z,)Y,X]
    (utilVoteInstances.py:263-271): with `numinst_threshs` = (t1, t2, ...) class i is taken
    where prob[i] > t_i, later classes overriding earlier ones, else 0; without them the
    arg-max class."""
    prob = np.squeeze(np.asarray(numinst_prob))
    if prob.ndim == 3:                      # 2-D data: [C,Y,X] -> [C,1,Y,X]
        prob = prob[:, None]
    threshs = kwargs.get('numinst_threshs')
    if not threshs:
        return np.argmax(prob, axis=0).astype(np.uint8)
    numinst = np.zeros(prob.shape[1:], dtype=np.uint8)
    for cls, th in enumerate(threshs, start=1):
        numinst[prob[cls] > th] = cls
    return numinst


def maybeLoadNuminst(f, **kwargs):
    if kwargs.get('numinst_key') is not None:
        return numinst_from_prob(np.array(f[kwargs['numinst_key']]), **kwargs)
    return None


def resolve_foreground(fg=None, numinst=None, mid=None, **kwargs):
    """the one foreground rule of the stage (utilVoteInstances.py:275-322): the
    `fg_key` array if that key is set, else `numinst > 0` if `numinst_key` is set,
    else the centre channel of the prediction; thresholded with `fg_thresh_vi` if
    positive, else `patch_threshold`.  Arrays are squeezed to the volume's own
    dimensions (the reference's file loader keeps a leading 1 in two of the three
    branches, which breaks its own 3-D non-blockwise path)."""
    if kwargs.get('fg_key') is not None and fg is not None:
        src = np.squeeze(np.asarray(fg))
    elif kwargs.get('numinst_key') is not None and numinst is not None:
        src = np.squeeze(np.asarray(numinst)) > 0
    else:
        assert mid is not None, "no foreground source"
        src = np.squeeze(np.asarray(mid))
    return src > getFgThreshold(**kwargs)


def loadFg(f, **kwargs):
    """utilVoteInstances.py:275-303 on an open container: (foreground bool, key)."""
    fg_key, numinst_key = kwargs.get('fg_key'), kwargs.get('numinst_key')
    if fg_key is not None:
        return resolve_foreground(fg=np.array(f[fg_key]), **kwargs), fg_key
    if numinst_key is not None:
        return resolve_foreground(numinst=maybeLoadNuminst(f, **kwargs), **kwargs), numinst_key
    mid = int(np.prod(kwargs['patchshape'])) // 2
    return resolve_foreground(mid=np.array(f[kwargs['aff_key']][mid]), **kwargs), \
        kwargs['aff_key']


def returnFg(affs, numinst, fg, **kwargs):
    """utilVoteInstances.py:306-322 on the arrays of one block."""
    mid = None
    if kwargs.get('fg_key') is None and kwargs.get('numinst_key') is None:
        mid = affs[int(np.prod(kwargs['patchshape'])) // 2]
    return resolve_foreground(fg=fg, numinst=numinst, mid=mid, **kwargs)


def getResKey(**kwargs):
    """suffix of the result datasets when `add_suffix` is set (utilVoteInstances.py:325-337):
    _<threshold digits>[_tfgc][_mws][_smp<sample digits>]."""
    digits = lambda v: str(v).replace('.', '')
    parts = [digits(kwargs['patch_threshold'])]
    if not kwargs.get('skipThinCover', False):
        parts.append('tfgc')                # thinned foreground cover
    if kwargs['mws']:
        parts.append('mws')
    if kwargs['sample'] < 1.0:
        parts.append('smp' + digits(kwargs['sample']))
    return '_' + '_'.join(parts)


def _open(aff_file):
    from .io_util import open_container
    return open_container(aff_file, 'r')


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
