"""Distribution of ink bounding boxes (grown by the stencil reach) on a synthetic workload: sizes the backward kernel's
box-sized shared-memory buffers.  numpy on the host; not the product path."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybpl_b200 import workload


def boxes(w, reach=12, gran=5):
    C = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pybpl_b200", "data", "basis_5_200.npy")).astype(np.float64)
    tso, sso = w["tok_stroke_off"].astype(np.int64), w["stroke_sub_off"].astype(np.int64)
    out = []
    for p in range(w["P"]):
        r0 = c0 = 1e9; r1 = c1 = -1e9
        for k in range(tso[p], tso[p + 1]):
            prev = w["position"][k].astype(np.float64)
            for s in range(sso[k], sso[k + 1]):
                X = C @ (w["invscale"][s] * w["ctrl"][s].astype(np.float64))
                X = X - (X[0] - prev)
                prev = X[-1]
                r, c = -X[:, 1], X[:, 0]
                m = (r >= 0) & (r <= 104) & (c >= 0) & (c <= 104)
                if m.any():
                    r0 = min(r0, np.floor(r[m]).min()); r1 = max(r1, np.ceil(r[m]).max())
                    c0 = min(c0, np.floor(c[m]).min()); c1 = max(c1, np.ceil(c[m]).max())
        if r1 < 0:
            out.append((0, 0)); continue
        h = min(104, r1 + reach) - max(0, r0 - reach) + 1
        wd = min(104, c1 + reach) - max(0, c0 - reach) + 1
        h = min(105, -(-int(h) // gran) * gran); wd = min(105, -(-int(wd) // gran) * gran)
        out.append((h, wd))
    return np.array(out)


if __name__ == "__main__":
    for name, w in (("bwd bench (1024, sigma 10)", workload.synth_tokens(1024, seed=2024, sigma_mode=10.0)),
                    ("cfg1 (4096)", workload.synth_tokens(2048, seed=1234, sigma_mode="uniform"))):
        b = boxes(w)
        area = b[:, 0] * (b[:, 1] | 1)
        print(name, "mean h,w", b.mean(0), "mean area", area.mean(), "frac of canvas", area.mean() / 11025)
        for cap in (3700, 4500, 5513, 6000, 7000, 8000, 8800, 9600, 11130):
            print(f"  cap {cap:6d}: {100.0 * (area <= cap).mean():5.1f} % fit")
        print("  percentiles area", np.percentile(area, [10, 25, 50, 75, 90, 95, 99]))
