#!/usr/bin/env python
"""Does the production batch (front end -> fix-up -> augmentation epilogue) gain from running the epilogue of batch k NEXT TO
the front-end grid of batch k + 1?  The front-end grid is bound by the SM (HBM 40 % busy), the epilogue by memory latency /
bandwidth; on one stream they only overlap in each other's tails, because six front-end CTAs leave no registers for an
epilogue CTA.  Here the front-end CTA's shared memory is padded so that only 5, 4 or 3 fit per SM and the epilogue runs on a second stream.

    python tools/cosched_probe.py [steps] -> one JSON object on stdout (profiles/r02_cosched_probe.json)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whisper_finetune_b200 as wft  # noqa: E402

B, NM, T = 64, 128, 3000
SEED, TM, FM, W = 42, 100, 43, 80


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = wft._lib.load()
    g = torch.Generator().manual_seed(1)
    n_sets = 4
    pcm = [(0.1 * torch.randn(B, 480000, generator=g)).clamp_(-1, 1).to(dev) for _ in range(n_sets)]
    plain = [torch.empty(B, NM, T, device=dev) for _ in range(n_sets)]
    n_out = 16   # an independent launch stays clear of every call since the last one that waited (ops._LAST_CALL)
    outs = [torch.empty(B, NM, T, device=dev) for _ in range(n_out)]
    sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev, priority=-1)
    res = {"batch": B, "steps": steps, "what": "ms per production batch (front end + fix-up + drawn augmentation epilogue), B = 64, "
           "4 rotating buffer sets; caps = front-end CTAs per SM"}

    def timed(body):
        body(0, 8)
        torch.cuda.synchronize()
        best = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            body(rep * steps, steps)
            e1.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1) / steps)
        best.sort()
        return best[len(best) // 2]

    # 1. what ships: one stream, three grids chained by programmatic dependent launches
    fe = wft.FrontEnd(n_mels=NM, device=dev, spec_augment=True, seed=SEED,
                      spec_augment_params={"time_mask_param": TM, "freq_mask_param": FM, "time_warp_w": W, "p": 1.0})

    def one_stream(first, n):
        for i in range(first, first + n):
            fe(pcm[i % n_sets], clip_offset=i * B, out=outs[i % n_out])

    wft.set_overlap(True)
    res["one_stream_ms"] = timed(one_stream)

    # 1b. the same batch cut into sub-batches whose un-warped features fit in the L2 (ncu: at B = 64 the epilogue re-reads 83 %
    # of the 98 MB the front end wrote from DRAM)
    def chunked(c):
        def body(first, n):
            for i in range(first, first + n):
                for a in range(0, B, c):
                    fe(pcm[i % n_sets][a:a + c], clip_offset=i * B + a, out=outs[i % n_out][a:a + c])
        return body

    res["one_stream_chunked_ms"] = {str(c): timed(chunked(c)) for c in (32,)}

    # 1c. the three-grid composition the fused call replaced (front end + fix-up grid, then the epilogue on finished features)
    plain3 = [torch.empty(B, NM, T, device=dev) for _ in range(n_out)]

    def three_grids(first, n):
        for i in range(first, first + n):
            wft.frontend_forward(pcm[i % n_sets], NM, out=plain3[i % n_out])
            torch.ops.wft.augment_drawn_out(plain3[i % n_out], SEED, i * B, TM, FM, W, 1.0, None, 0.0, False, outs[i % n_out])

    res["one_stream_three_grids_ms"] = timed(three_grids)

    # 2. front end alone / epilogue alone at the same caps (what each costs when it has the GPU to itself)
    def fe_only(first, n):
        for i in range(first, first + n):
            wft.frontend_forward(pcm[i % n_sets], NM, out=plain[i % n_sets])

    def aug_only(first, n):
        for i in range(first, first + n):
            torch.ops.wft.augment_drawn_out(plain[i % n_sets], SEED, i * B, TM, FM, W, 1.0, None, 0.0, False, outs[i % n_out])

    res["epilogue_alone_ms"] = timed(aug_only)

    # 3. two streams: A = front end + fix-up (independent batches), B = epilogue of the batch A just finished
    def two_streams(first, n):
        cur = torch.cuda.current_stream(dev)
        sa.wait_stream(cur)
        sb.wait_stream(cur)
        done_a, done_b = {}, {}
        for i in range(first, first + n):
            s = i % n_sets
            with torch.cuda.stream(sa):
                if i - n_sets in done_b:            # plain[s] is free once the epilogue that read it is done
                    sa.wait_event(done_b[i - n_sets])
                wft.frontend_forward(pcm[s], NM, out=plain[s])
                ev = torch.cuda.Event()
                ev.record(sa)
                done_a[i] = ev
            with torch.cuda.stream(sb):
                sb.wait_event(done_a[i])
                torch.ops.wft.augment_drawn_out(plain[s], SEED, i * B, TM, FM, W, 1.0, None, 0.0, False, outs[i % n_out])
                ev = torch.cuda.Event()
                ev.record(sb)
                done_b[i] = ev
        cur.wait_stream(sa)
        cur.wait_stream(sb)

    # the cap that matters is CTAs PER SM (independent batches overlap, so a smaller grid alone just lets the next batch's CTAs
    # in): pad the front-end CTA's shared memory so that only 5 / 4 / 3 fit
    res["caps"] = {}
    for per_sm, extra in ((6, 0), (4, 16 * 1024)):
        lib.wft_debug_set_extra_smem(extra)
        r = {"front_end_alone_ms": timed(fe_only), "two_streams_ms": timed(two_streams), "one_stream_ms": timed(one_stream)}
        res["caps"][str(per_sm)] = r
    lib.wft_debug_set_extra_smem(0)
    wft.set_overlap(False)
    # the two-stream result must equal the one-stream result
    torch.cuda.synchronize()
    ref = fe(pcm[0], clip_offset=0).clone()
    lib.wft_debug_set_extra_smem(6 * 1024)
    wft.set_overlap(True)
    two_streams(0, 4)
    torch.cuda.synchronize()
    lib.wft_debug_set_extra_smem(0)
    res["two_streams_equal_one_stream"] = bool(torch.equal(ref, outs[0]))
    res["clips_per_s"] = {"one_stream": B / res["one_stream_ms"] * 1e3,
                          **{f"two_streams_cap_{k}": B / v["two_streams_ms"] * 1e3 for k, v in res["caps"].items()}}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
