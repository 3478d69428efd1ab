#!/usr/bin/env python3
"""Largest deviation of the device's SVC Material.response from the reference fixture (tests/golden/svc_hill.npz), per
output, next to the oracle's: the number behind the 2e-6 tolerance of tests/test_gpu_parity.py."""
import os
import sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np  # noqa: E402
from helpers import plf_material_from_golden  # noqa: E402
from pylabfea_b200 import _lib  # noqa: E402
from oracle import oracle as orc  # noqa: E402

g = dict(np.load(os.path.join(R, 'tests', 'golden', 'svc_hill.npz')))
ctx = _lib.Context(0)
for tag in ('pe', 'ps'):
    ctx.set_materials([plf_material_from_golden(g)], planestress=(tag == 'ps'))
    dev = ctx.response_batch(g[tag + '_sig'], g[tag + '_epl'], g[tag + '_deps'])
    ora = orc.response_batch([orc.Mat.from_golden(g)], None, g[tag + '_sig'], g[tag + '_epl'], g[tag + '_deps'], g[tag + '_CV'][None])
    for name, out in (('device', dev), ('oracle', ora)):
        fy, so, dp, ct, ns = out
        e_sig = max(np.max(np.abs(so[i] - g[tag + '_sig_out'][i])) / np.max(np.abs(g[tag + '_sig_out'][i])) for i in range(len(fy)))
        e_ct = max(np.max(np.abs(ct[i] - g[tag + '_ctan'][i])) / np.max(np.abs(g[tag + '_ctan'][i])) for i in range(len(fy)))
        e_dp = max(np.max(np.abs(dp[i] - g[tag + '_depl'][i])) / max(np.max(np.abs(g[tag + '_depl'][i])), 1e-4) for i in range(len(fy)))
        e_fy = np.max(np.abs(fy - g[tag + '_fy'])) / float(g['sy'])
        print('%s %s: n=%d nsteps equal=%s  fy %.2e  sig %.2e  depl %.2e  ctan %.2e' %
              (tag, name, len(fy), np.array_equal(ns, g[tag + '_nsteps']), e_fy, e_sig, e_dp, e_ct))
