"""GPU-resident persistence-image feature table and its on-disk cache.

Reference behaviour being replaced:
  * loaddatas.py:56-103 `compute_persistence_image`: cache file './data/TLCGNN/<Name>.npy' holding
    pi_sg float64[E, 25]; rows ordered (train_pos, train_neg, val_pos, val_neg, test_pos, test_neg) (:65-66);
  * baselines/TLCGNN.py:35-53 `decode`: every call slices / fancy-indexes the HOST table, converts to float32 and
    uploads (`torch.Tensor(PI.reshape(len(total_edges), -1)).cuda()`).

Here the float64 table is uploaded once (or never leaves the GPU: `tlc_vicinity_pi_dev` leaves the rows in HBM) and
`rows()` gathers + rounds the step's rows on the device (kernel 5, `tlc_pi_gather`).  torch holds the device memory;
only data pointers cross the C-ABI.
"""
import os

import numpy as np

from . import _lib as L

SPLIT_NAMES = ("train_pos", "train_neg", "val_pos", "val_neg", "test_pos", "test_neg")


def cache_filename(data_name, cache_dir="./data/TLCGNN"):
    """loaddatas.py:57-62 (photo -> Photo, computers -> Computers)."""
    if data_name == "photo":
        data_name = "Photo"
    if data_name == "computers":
        data_name = "Computers"
    return os.path.join(cache_dir, data_name + ".npy")


class PITable:
    def __init__(self, pi_sg, splits=None, device=0):
        """pi_sg: float64[E, r2] numpy array (as np.load of the cache returns) or a CUDA float64 tensor;
        splits: the six lengths (train_pos, train_neg, val_pos, val_neg, test_pos, test_neg) of loaddatas.py:67-69."""
        import torch
        self.device = int(device)
        if isinstance(pi_sg, np.ndarray):
            a = np.ascontiguousarray(pi_sg, dtype=np.float64)
            self.table = torch.from_numpy(a).to(torch.device("cuda", self.device))
        else:
            assert pi_sg.is_cuda and pi_sg.dtype == torch.float64
            self.table = pi_sg.contiguous()
            self.device = self.table.device.index
        assert self.table.dim() == 2
        self.rows_total, self.r2 = int(self.table.shape[0]), int(self.table.shape[1])
        self.splits = None
        if splits is not None:
            s = [int(x) for x in splits]
            assert len(s) == 6 and sum(s) == self.rows_total, "splits must cover the table"
            self.splits = dict(zip(SPLIT_NAMES, s))

    # ---- the .npy cache (loaddatas.py:62-64,102) ----
    @classmethod
    def from_npy(cls, filename, splits=None, device=0):
        return cls(np.load(filename), splits=splits, device=device)

    def save_npy(self, filename):
        os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
        np.save(filename, self.table.cpu().numpy())   # float64[E, r2], what np.save(filename, pi.pi_sg) writes

    # ---- device gathers ----
    def gather(self, index=None, start=0, n=None, out=None):
        """float32[n, r2] CUDA tensor: table[index] (int64 CUDA tensor / array) or table[start:start+n]."""
        import torch
        dev = self.table.device
        idx_ptr = None
        if index is not None:
            if not isinstance(index, torch.Tensor):
                index = torch.as_tensor(np.asarray(index, dtype=np.int64))
            index = index.to(device=dev, dtype=torch.int64).contiguous()
            n = int(index.numel())
            idx_ptr = index.data_ptr()
        elif n is None:
            n = self.rows_total - start
        if out is None:
            out = torch.empty((n, self.r2), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        L.check(L.lib().tlc_pi_gather(self.device, self.table.data_ptr(), self.rows_total, self.r2, idx_ptr, int(start),
                                      int(n), out.data_ptr(), st))
        return out

    def rows(self, kind="train", index=None):
        """the rows baselines/TLCGNN.py:35-53 selects, as the float32 CUDA tensor `new_x` it builds.
        kind='train': `index` = np.random.randint(0, train_neg, train_pos) of :38 (drawn by the caller, as there)."""
        import torch
        s = self.splits
        assert s is not None, "splits needed"
        tp, tn, vp, vn = s["train_pos"], s["train_neg"], s["val_pos"], s["val_neg"]
        if kind == "train":
            dev = self.table.device
            idx = torch.as_tensor(np.asarray(index, dtype=np.int64)) if not isinstance(index, torch.Tensor) else index
            idx = idx.to(device=dev, dtype=torch.int64)
            # numpy semantics of PI[tp:tp+tn][index]: relative to the slice, negative values wrap, else IndexError
            if idx.numel() and (int(idx.min()) < -tn or int(idx.max()) >= tn):
                raise IndexError("index out of bounds for the train_neg slice")
            idx = torch.where(idx < 0, idx + tn, idx)
            full = torch.cat([torch.arange(tp, device=dev, dtype=torch.int64), idx + tp])
            return self.gather(index=full)
        if kind == "val":
            return self.gather(start=tp + tn, n=vp + vn)
        if kind == "test":
            return self.gather(start=tp + tn + vp + vn, n=self.rows_total - (tp + tn + vp + vn))
        raise ValueError(kind)
