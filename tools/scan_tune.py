#!/usr/bin/env python
"""Tuning aid: builds bench.py's workload once, then times the scan stage (K1 / K1b) under several environment settings.
usage: python tools/scan_tune.py [workload] 'NAME=VAL,NAME=VAL' 'NAME=VAL' ...   ('' = defaults)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import ntedit_b200 as nb  # noqa: E402


def main():
    args = sys.argv[1:]
    workload = "3Gbp_k25_4GiB_m1"
    if args and args[0] in bench.WORKLOADS:
        workload = args.pop(0)
    w = bench.WORKLOADS[workload]
    nb.lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    filt = torch.zeros(w["fbytes"] + 64, dtype=torch.uint8, device=dev)
    bloom = nb.BloomFilter.wrap_device(filt.data_ptr(), w["fbytes"], w["k"], w["h"], counting=bool(w.get("counting")), device=0)
    buf, offs = bench.build_workload(w, dev, 0, bloom, nb)
    bench.finish_filter(w, filt, dev)
    torch.cuda.synchronize()
    batch = nb.Batch.wrap_device(buf.data_ptr(), offs, device=0)
    params = nb.default_params(mode=w["mode"], snv=int(w.get("snv", 0)))
    base_env = dict(os.environ)
    for cfg in (args or [""]):
        os.environ.clear()
        os.environ.update(base_env)
        for kv in [x for x in cfg.split(",") if x]:
            k, v = kv.split("=")
            os.environ[k] = v
        params.segment_len = int(os.environ.get("NTB_TUNE_SEGMENT_LEN", "0"))
        out = []
        walls = []
        import time
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = nb.kmerize_and_correct_device(batch, bloom, params, host_buf=None)
            walls.append(round(1000 * (time.perf_counter() - t0), 1))
            st = res.stats().as_dict()
            res.free()
            out.append(st)
        st = out[-1]
        if os.environ.get("NTB_TUNE_E2E"):
            import time
            host = torch.empty(len(buf), dtype=torch.uint8, pin_memory=True)
            work = torch.empty(len(buf), dtype=torch.uint8, pin_memory=True)
            host.copy_(buf)
            torch.cuda.synchronize()
            for _ in range(3):
                work.copy_(host)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                res = nb.kmerize_and_correct(work.numpy(), offs, bloom, params)
                t2 = time.perf_counter()
                d = res.stats().as_dict()
                t3 = time.perf_counter()
                res.free()
                t4 = time.perf_counter()
                print(json.dumps({"e2e_call_ms": round(1000 * (t2 - t1), 1), "free_ms": round(1000 * (t4 - t3), 1),
                                  "h2d": round(d["ms_h2d"], 1), "scan": round(d["ms_scan"], 1), "walk": round(d["ms_walk"], 1),
                                  "host": round(d["ms_host"], 1), "d2h": round(d["ms_d2h"], 1)}), flush=True)
        print(json.dumps({"env": cfg, "wall_ms": walls, "ms_scan": [round(o["ms_scan"], 2) for o in out], "ms_pre": round(st["ms_pre"], 2), "ms_walk": round(st["ms_walk"], 2), "ms_host": round(st["ms_host"], 2),
                          "launches": st["kernel_launches"], "sites": st["sites"], "edits": st["edits"]}), flush=True)


if __name__ == "__main__":
    main()
