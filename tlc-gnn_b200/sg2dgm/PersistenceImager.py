"""Mirror of sg2dgm/PersistenceImager.pyx: PersistenceImager(resolution).transform(dgm, skew=True)
(PersistenceImager.pyx:207-242, 352-388) for the isotropic sigma = 1 kernel and linear_ramp weights that
the hot path uses; runs kernel 4.  The general bivariate-normal kernel, `fit` and the setters of the
reference class are never reached by that path and are out of scope.
"""
import numpy as np

from tlc_b200 import api


class PersistenceImager:
    def __init__(self, birth_range=None, pers_range=None, pixel_size=None, resolution=5, weight=None,
                 weight_params=None, kernel=None, kernel_params=None, device=0):
        if birth_range not in (None, (0.0, 1.0), (0, 1)) or pers_range not in (None, (0.0, 1.0), (0, 1)) \
                or pixel_size is not None or weight is not None or kernel is not None \
                or weight_params not in (None, {}) or kernel_params is not None:
            raise NotImplementedError("only the defaults the TLC-GNN path uses are implemented: ranges (0,1), "
                                      "linear_ramp weight, isotropic unit Gaussian (PersistenceImager.pyx:220-230)")
        self._resolution = (resolution, resolution)
        self._pixel_size = 1.0 / resolution
        self._birth_range = (0.0, 1.0)
        self._pers_range = (0.0, 1.0)
        self.device = device
        # _create_mesh  PersistenceImager.pyx:311-314
        self._bpnts = np.linspace(0.0, 1.0 + self._pixel_size, resolution + 1, endpoint=False, dtype=np.float64)
        self._ppnts = self._bpnts.copy()

    @property
    def resolution(self):
        return self._resolution

    def transform(self, pers_dgm, skew=True):
        d = np.array(pers_dgm, dtype=np.float64).reshape(-1, 2)
        if not skew:  # already (birth, persistence): the kernel skews, so un-skew first
            d = d.copy()
            d[:, 1] = d[:, 0] + d[:, 1]
        return api.pimg_transform(d, self._resolution[0], device=self.device)
