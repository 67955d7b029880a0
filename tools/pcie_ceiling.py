"""What the host link of this box sustains for the e2e leg's traffic pattern (bench.py `e2e`: 80.3 MB up + 80.3 MB down per
matvec): pinned-memory H2D alone, D2H alone, and both at once on two streams.  Prints one JSON line.
    python tools/pcie_ceiling.py [bytes] [reps]"""
import json
import sys

import torch


def main():
    nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 80314576
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    n = nbytes // 8
    dev = torch.device("cuda:0")
    h_in = torch.empty(n, dtype=torch.float64).pin_memory()
    h_out = torch.empty(n, dtype=torch.float64).pin_memory()
    h_in.uniform_()
    d_in = torch.empty(n, dtype=torch.float64, device=dev)
    d_out = torch.rand(n, dtype=torch.float64, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def h2d():
        d_in.copy_(h_in, non_blocking=True)

    def d2h():
        h_out.copy_(d_out, non_blocking=True)

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    t_h2d, t_d2h, t_both = timed(h2d), timed(d2h), timed(both)
    gb = nbytes / 1e9
    print(json.dumps({"bytes_per_direction": nbytes, "reps": reps,
                      "h2d_alone_ms": t_h2d, "h2d_alone_gbs": gb / (t_h2d * 1e-3),
                      "d2h_alone_ms": t_d2h, "d2h_alone_gbs": gb / (t_d2h * 1e-3),
                      "both_ms": t_both, "both_gbs_per_direction": gb / (t_both * 1e-3),
                      "note": "both = one H2D and one D2H of the same size in flight at once: the floor of a pipelined "
                              "host-buffer matvec is both_ms (+ the first chunk's upload and the last chunk's download)"}))


if __name__ == "__main__":
    main()
