"""CPU tests of the N>1 host logic: contiguous frame sharding and the rank-0 gather, run with two
gloo ranks (no GPU; the per-frame work is a stand-in computed with the oracle)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_range_partitions_exactly():
    from image_b200.shard import frame_range
    for n in (0, 1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            spans = [frame_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_gloo_ranks_shard_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent("""
        import os, sys, json
        import numpy as np
        sys.path.insert(0, %r)
        import torch.distributed as dist
        from image_b200 import synth
        from image_b200.shard import run_sharded
        from oracle import pyoracle as po
        dist.init_process_group("gloo")
        frames = np.stack([synth.frame_shapes(10 + i, 48, 64) for i in range(5)])
        def work(fs):
            return [int(po.canny(f)[1]) for f in fs]          # per-frame edge-pixel count
        res = run_sharded(work, frames)
        if dist.get_rank() == 0:
            json.dump(res, open(%r, "w"))
        dist.barrier()
        dist.destroy_process_group()
    """ % (ROOT, str(tmp_path / "out.json"))))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env, timeout=240)
    import json
    from image_b200 import synth
    from oracle import pyoracle as po
    got = json.load(open(tmp_path / "out.json"))
    want = [int(po.canny(synth.frame_shapes(10 + i, 48, 64))[1]) for i in range(5)]
    assert got == want


def test_rshim_sources_compile_against_rcpp_interface():
    """The four replacement bodies of the reference's Rcpp exports (image_b200/rshim) must compile
    against include/b2f.h with an Rcpp-shaped header (R/Rcpp are absent here: oracle/stubs mock)."""
    for f in sorted(os.listdir(os.path.join(ROOT, "image_b200", "rshim"))):
        if f.endswith(".cpp"):
            subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "oracle", "stubs"),
                                   "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "image_b200", "rshim", f)])


def test_gpu_numa_cpus_from_a_fake_sysfs(tmp_path):
    from image_b200.shard import gpu_numa_cpus, _parse_cpulist
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices/system/node/node1"
    node.mkdir(parents=True)
    (node / "cpulist").write_text("32-63\n")
    assert gpu_numa_cpus("00000000:1B:00.0", sysfs=str(tmp_path)) == set(range(32, 64))
    (dev / "numa_node").write_text("-1\n")
    assert gpu_numa_cpus("0000:1b:00.0", sysfs=str(tmp_path)) is None
    assert gpu_numa_cpus("0000:ff:00.0", sysfs=str(tmp_path)) is None
