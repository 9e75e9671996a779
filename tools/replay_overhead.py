"""Host-side / stream-ordering overhead of the library's CUDA-graph replay (gnb_model_forward): cfg4 forwards back to back on the
legacy default stream (fork / join through the library's own stream) and on a torch side stream (direct replay), with and
without the cross-context chain event:  python tools/replay_overhead.py [graphs]"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 2 and sys.argv[2] == "child":
    import torch                     # noqa: E402
    import graphnets_b200 as gn      # noqa: E402
    import workloads as W            # noqa: E402
    from bench import synth          # noqa: E402
    B = int(sys.argv[1])
    side = os.environ.get("RO_SIDE") == "1"
    adj, ef, nf = synth("cfg4", B, 1000)
    model = W.to_gn_model(gn, W.model_params("cfg4"))
    stream = torch.cuda.Stream() if side else torch.cuda.current_stream()
    with torch.cuda.stream(stream):
        x = gn.batch_compact(adj, ef, nf)
        for _ in range(6):
            model(x, precision="auto")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        e0.record()
        for _ in range(n):
            model(x, precision="auto")
        e1.record()
        torch.cuda.synchronize()
    print("side_stream=%d chain=%s graph=%s: %.3f ms per forward" % (side, os.environ.get("GNB_CHAIN_FORWARDS", "1"), os.environ.get("GNB_CUDA_GRAPH", "1"),
                                                                   e0.elapsed_time(e1) / n))
else:
    B = sys.argv[1] if len(sys.argv) > 1 else "4096"
    for side in ("0", "1"):
        for chain in ("1", "0"):
            env = dict(os.environ, RO_SIDE=side, GNB_CHAIN_FORWARDS=chain)
            subprocess.run([sys.executable, __file__, B, "child"], env=env)
    subprocess.run([sys.executable, __file__, B, "child"], env=dict(os.environ, RO_SIDE="0", GNB_CUDA_GRAPH="0"))
