#!/usr/bin/env python
"""Generates tests/golden/tap_channel_stats.json from the reference's shipped feature-map data sets (run in the build
container, where /root/reference exists): per channel, the fraction of rows in which the channel is non-zero, averaged over
the data sets with more than 500 rows, for the 88-channel (re_lu_10) and 96-channel (re_lu_15) taps.  The data sets are real
outputs of the reference's backbone on face crops, so the statistics pin which channels its trained
backbone uses and which are dead."""
import glob
import json
import os

import numpy as np

REF = "/root/reference/FeatureMaps-Datasets"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tap_channel_stats.json")
res = {}
for c, pat in ((88, "*_88_0.7_1.npz"), (96, "*_96_0.7_1.npz")):
    fr, files = [], []
    for f in sorted(glob.glob(os.path.join(REF, pat))):
        x = np.load(f)["features"]
        if len(x) > 500:
            fr.append((x > 0).mean(0))
            files.append({"file": os.path.basename(f), "rows": int(len(x)), "dead": np.where((x == 0).all(0))[0].tolist()})
    res[str(c)] = {"nonzero_fraction": [round(float(v), 6) for v in np.mean(fr, 0)], "files": files}
with open(OUT, "w") as fh:
    json.dump(res, fh)
print("wrote", OUT, {k: len(v["nonzero_fraction"]) for k, v in res.items()})
