"""reference evaluation_tts.py:50-97 ``gen_parameters`` re-typed with the modules it uses passed in (the original imports
pyworld / pysptk / an HTS question set at module load and cannot travel to the GPU box).  TEST INFRASTRUCTURE: pinned to the
unmodified function by tests/golden/eval.npz (tests/golden/make_golden.py::gen_eval executes the original where it lies).
``paramgen`` needs ``mlpg(mean_frames, variance_frames, windows)``, ``P`` needs ``inv_scale(x, mean, std)``."""
import numpy as np


def gen_parameters(y_predicted, Y_mean, Y_std, mge_training, stream_sizes, windows, paramgen, P):
    """Y_mean / Y_std: the acoustic statistics, (187,) arrays (the original indexes dicts with ["acoustic"])."""
    mgc_dim, lf0_dim, vuv_dim, bap_dim = stream_sizes
    nw = len(windows)
    lf0_0 = mgc_dim
    vuv_0 = lf0_0 + lf0_dim
    bap_0 = vuv_0 + vuv_dim
    if mge_training:                                                       # :62-81
        mgc, lf0 = y_predicted[:, :lf0_0], y_predicted[:, lf0_0:vuv_0]
        vuv, bap = y_predicted[:, vuv_0], y_predicted[:, bap_0:]
        mgc = paramgen.mlpg(mgc, np.ones(mgc.shape[-1]), windows)          # MLPG on normalised features, unit variance
        lf0 = paramgen.mlpg(lf0, np.ones(lf0.shape[-1]), windows)
        bap = paramgen.mlpg(bap, np.ones(bap.shape[-1]), windows)
        mgc = P.inv_scale(mgc, Y_mean[:mgc_dim // nw], Y_std[:mgc_dim // nw])
        lf0 = P.inv_scale(lf0, Y_mean[lf0_0:lf0_0 + lf0_dim // nw], Y_std[lf0_0:lf0_0 + lf0_dim // nw])
        bap = P.inv_scale(bap, Y_mean[bap_0:bap_0 + bap_dim // nw], Y_std[bap_0:bap_0 + bap_dim // nw])
        vuv = P.inv_scale(vuv, Y_mean[vuv_0], Y_std[vuv_0])
    else:                                                                  # :82-95
        y_predicted = P.inv_scale(y_predicted, Y_mean, Y_std)
        mgc, lf0 = y_predicted[:, :lf0_0], y_predicted[:, lf0_0:vuv_0]
        vuv, bap = y_predicted[:, vuv_0], y_predicted[:, bap_0:]
        Y_var = Y_std * Y_std
        mgc = paramgen.mlpg(mgc, Y_var[:lf0_0], windows)                   # MLPG with the real variances
        lf0 = paramgen.mlpg(lf0, Y_var[lf0_0:vuv_0], windows)
        bap = paramgen.mlpg(bap, Y_var[bap_0:], windows)
    return mgc, lf0, vuv, bap
