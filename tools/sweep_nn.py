#!/usr/bin/env python
"""Time every nn_kernel launch variant (URED_NN_VARIANT / URED_NN_NSPLIT) on the BASELINE shapes and check that each one
returns the same bits as the unmodified reference op (oracle/_ref) -- or, for shapes the reference op is slow on, as
variant 0.  Development tool: writes one JSON line per (shape, variant, nsplit) to stdout / --out.

    python tools/sweep_nn.py [--out gpurun_out/sweep_nn.jsonl] [--shapes cfg2,cfg1,r125,cfg4] [--variants 0,1,2]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {
    # name: (pairs, n1, n2)
    "cfg2": (640, 2048, 2048),
    "cfg1": (32, 2048, 2048),
    "r125": (125, 2048, 2048),   # one rank's share of cfg3 (1000 shapes over 8 GPUs)
    "cfg4": (16, 16384, 16384),
    "odd": (48, 2000, 1000),     # the reference unit test's timing shape (unit_test.py:39-40)
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--shapes", default="cfg2,cfg1,r125,cfg4,odd")
    ap.add_argument("--variants", default="0,1,2,3,4,5,6,7")
    ap.add_argument("--nsplits", default="1,2,4,8")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()

    import torch
    import ured_b200 as ured
    from bench import synth
    lib = ured._native.load()
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = open(args.out, "w") if args.out else None

    ref_op = None
    try:
        from oracle import build as obuild
        import importlib.util
        path = obuild.ref_module_path()
        if path:
            spec = importlib.util.spec_from_file_location(obuild.REF_NAME, path)
            ref_op = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(ref_op)
    except Exception as exc:  # pragma: no cover
        print("reference op unavailable:", exc, file=sys.stderr)

    for name in args.shapes.split(","):
        B, n1, n2 = SHAPES[name]
        x, gt = synth(B, n2, n1, seed=31)
        c1, c2 = gt.to(dev), x.to(dev)
        pk1, pk2 = ured.PackedClouds(c1), ured.PackedClouds(c2)
        d1 = torch.empty(B, n1, device=dev); d2 = torch.empty(B, n2, device=dev)
        i1 = torch.empty(B, n1, device=dev, dtype=torch.int32); i2 = torch.empty(B, n2, device=dev, dtype=torch.int32)
        stream = torch.cuda.current_stream(dev).cuda_stream
        want = None
        if ref_op is not None and name != "cfg4":
            w = [torch.zeros(B, n1, device=dev), torch.zeros(B, n2, device=dev),
                 torch.zeros(B, n1, device=dev, dtype=torch.int32), torch.zeros(B, n2, device=dev, dtype=torch.int32)]
            ref_op.forward(c1, c2, *w)
            torch.cuda.synchronize()
            want = w
        for v in [int(t) for t in args.variants.split(",")]:
            for ns in [int(t) for t in args.nsplits.split(",")]:
                if ns > min(n1, n2) // 512:
                    continue
                os.environ["URED_NN_VARIANT"] = str(v)
                os.environ["URED_NN_NSPLIT"] = str(ns)
                sb = lib.ured_nn_scratch_bytes(B, n1, n2)
                scratch = torch.empty(max(sb, 256), dtype=torch.uint8, device=dev)

                def run():
                    rc = lib.ured_nn_packed(c1.data_ptr(), pk1.packed.data_ptr(), n1, c2.data_ptr(), pk2.packed.data_ptr(), n2,
                                            B, 1, B, None, None, d1.data_ptr(), d2.data_ptr(), i1.data_ptr(), i2.data_ptr(),
                                            scratch.data_ptr(), sb, 0, stream)
                    ured._native.check(rc, "ured_nn_packed")

                for _ in range(3):
                    run()
                torch.cuda.synchronize()
                times = []
                for _ in range(args.reps):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); run(); e1.record()
                    torch.cuda.synchronize()
                    times.append(e0.elapsed_time(e1))
                times.sort()
                med = times[len(times) // 2]
                got = [d1, d2, i1, i2]
                if want is None:
                    want = [t.clone() for t in got]   # first variant is the yardstick when the reference op is not used
                    same = None
                else:
                    same = all(torch.equal(g, w) for g, w in zip(got, want))
                rec = {"shape": name, "B": B, "n1": n1, "n2": n2, "variant": v, "nsplit": ns, "ms_median": med, "ms_min": times[0],
                       "tpair_s": 2.0 * B * n1 * n2 / (med * 1e-3) / 1e12, "bit_exact": same,
                       "yardstick": "reference op" if (ref_op is not None and name != "cfg4") else "first variant"}
                line = json.dumps(rec)
                print(line, flush=True)
                if out:
                    out.write(line + "\n"); out.flush()
    os.environ.pop("URED_NN_VARIANT", None)
    os.environ.pop("URED_NN_NSPLIT", None)


if __name__ == "__main__":
    main()
