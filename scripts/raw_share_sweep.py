"""acwm_search_host end to end (pinned host text -> host count + positions) against ACWM_HOST_RAW_PERCENT, the share of
the text that travels unpacked beside the host-packed rest.  One JSON line per (size, share)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acwm_pkg  # noqa: E402

acwm = acwm_pkg.load()
dg = acwm_pkg.submodule("datagen")


def main():
    import torch
    base = dg.text_host(1 << 30, 4, 1)
    pats = dg.patterns_with_hits(base, 1000, 16, 4, 2)
    mt = acwm.Matcher(acwm.WM, pats, 4).upload(device=0)
    for mib in (128, 1024):
        text = torch.from_numpy(base[: mib << 20]).pin_memory()
        for share in ("default", 0, 8, 15, 22, 30, 40):
            if share == "default":
                os.environ.pop("ACWM_HOST_RAW_PERCENT", None)
            else:
                os.environ["ACWM_HOST_RAW_PERCENT"] = str(share)
            ts = []
            for rep in range(7):
                t = time.perf_counter()
                count, pos = mt.search_host(text, cap=1 << 20)
                ts.append(time.perf_counter() - t)
            ts = sorted(ts[1:])
            print(json.dumps({"cores": os.cpu_count(), "text_mib": mib, "raw_percent": share, "best_GBps": text.numel() / ts[0] / 1e9,
                              "median_GBps": text.numel() / ts[len(ts) // 2] / 1e9, "h2d_bytes": mt.last_h2d_bytes,
                              "count": count}), flush=True)


if __name__ == "__main__":
    main()
