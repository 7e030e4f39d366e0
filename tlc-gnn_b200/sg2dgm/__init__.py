"""Drop-in mirror of the reference's `sg2dgm` package (pkuyzy/TLC-GNN), backed by libtlc_b200.so.

Same module names, function names, argument meaning and error behaviour as
sg2dgm/riccidist2dgm.py, sg2dgm/accelerated_PD.py and sg2dgm/PersistenceImager.pyx; the work is
done by hand-written sm_100a CUDA kernels through the C-ABI in include/tlc_b200.h.  Put the
directory `tlc-gnn_b200/` on sys.path in place of the reference root and `loaddatas.py` /
`baselines/TLCGNN.py` / the KD generators run unchanged.
"""
