"""The multi-GPU pipeline behind the C ABI (sccd_comm_create / sccd_ccd_sharded, csrc/shard.cu).

On the single test GPU: a communicator of world 1 runs the whole sliced path -- slice boxes from
the replicated vertex boxes, 64-bit (key, index) records, exchange (a local copy), record sort,
exact boxes REBUILT by the gather -- and must reproduce the plain pipeline bit for bit.
With >= 2 GPUs visible, tests/mgpu_check.py is launched under torchrun (one process per GPU,
NCCL): same TOI as one GPU, the ranks' pair lists partition the single-GPU list."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("which,f32,axis", [("small", False, 0), ("c1", False, 0), ("c1", True, 0),
                                            ("c1", False, 1), ("pile", False, 0)])
def test_sliced_build_on_one_rank_is_the_plain_pipeline(sccd, scene_small, scene_c1, which, f32, axis):
    s = sccd.scenes.blob_pile(300, seed=5) if which == "pile" else \
        {"small": scene_small, "c1": scene_c1}[which]
    c = sccd.Context(0)
    try:
        if f32:
            c.set_scalar_type(sccd.capi.F32)
        c.set_option(sccd.capi.OPT_SWEEP_AXIS, axis)
        c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
        toi = c.ccd()
        st0 = c.stats()
        lists = [c.broad_phase(0), c.broad_phase(1)]
        c.comm_create(None, 0, 1)
        assert c.ccd_sharded() == toi
        st1 = c.stats()
        assert st1["n_pairs"] == st0["n_pairs"] and st1["n_records"] == st0["n_records"]
        assert st1["grid_cells"] == st0["grid_cells"]
        for k in (0, 1):     # the lists the sharded call swept: same pairs in the same order
            assert np.array_equal(c.broad_phase(k), lists[k])
        with pytest.raises(sccd.SccdError):
            c.get_boxes()                     # a rank only holds its slice
        # host-buffer entry (this rank copies 1 / world of the mesh)
        assert c.ccd_sharded_host(s["V0"], s["V1"], s["E"], s["F"]) == toi
        c.comm_destroy()
        assert c.ccd() == toi                 # and back to the plain pipeline
    finally:
        c.close()


def test_sharded_call_without_communicator_is_a_state_error(sccd, scene_small):
    s = scene_small
    c = sccd.Context(0)
    try:
        c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
        with pytest.raises(sccd.SccdError) as e:
            c.ccd_sharded()
        assert e.value.code == sccd.capi.ERR_STATE
        with pytest.raises(sccd.SccdError):
            c.comm_create(None, 0, 2)         # world > 1 needs rank 0's id
    finally:
        c.close()


def test_sliced_build_reports_bad_indices(sccd, scene_small):
    s = scene_small
    c = sccd.Context(0)
    try:
        F = s["F"].copy(order="F")
        F[3, 1] = 10 ** 6
        c.upload_mesh(s["V0"], s["V1"], s["E"], F)
        c.comm_create(None, 0, 1)
        with pytest.raises(sccd.SccdError) as e:
            c.ccd_sharded()
        assert e.value.code == sccd.capi.ERR_ARG
    finally:
        c.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_partition_the_problem_under_nccl(world):
    """One process per GPU under torchrun (skipped when fewer GPUs are visible): the C-ABI
    sharded pipeline returns the single-GPU TOI on configs 1 and 3-shape, and the ranks' pair
    lists concatenate to the single-GPU list."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    for workload in ("c1", "pile"):
        p = subprocess.run(
            [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
             "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
             os.path.join(ROOT, "tests", "mgpu_check.py"), workload, "--quick"],
            capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
        assert '"ok": true' in p.stdout
