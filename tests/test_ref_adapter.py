"""The drop-in classes against traccc's real types.

oracle/_ref/libtraccc_ref_adapter.so = include/traccc_b200/traccc_adapter.hpp compiled with nvcc
against the reference's own headers (edm::spacepoint_collection / seed_collection /
measurement_collection, algorithm<>, memory_resource, cuda::stream_wrapper,
bound_track_parameters_collection_types) and the vecmem stand-in, run in the call sequence of
examples/run/cuda/apps/seeding_example_cuda.cpp next to the reference's own
traccc::cuda::triplet_seeding_algorithm. Checked with the semantics of the reference's
soa_comparator (order-free seed sets): ours == CPU reference (also in order), the reference's
CUDA code >= 99.9 %; parameters read back through bound_track_parameters' accessors within 1e-5
of the reference's host parameter estimation."""
import numpy as np
import pytest

from oracle import oracle
from tests.helpers import rel_close
from traccc_b200 import toy_detector

pytestmark = pytest.mark.gpu


def _set(s):
    return set(zip(s["bottom"].tolist(), s["middle"].tolist(), s["top"].tolist()))


@pytest.mark.parametrize("n,seed,resizable", [(300, 7, False), (2000, 8, True), (10000, 9, False)])
def test_adapter_with_reference_types(n, seed, resizable):
    ev = toy_detector.generate_event(n, seed)
    assert oracle.ref_adapter_lib() is not None, "oracle/_ref/libtraccc_ref_adapter.so is missing"
    ours = oracle.ref_adapter_run(1, ev, resizable_input=resizable)
    cpu = oracle.ref_run(ev.xyz, ev.var_z, ev.var_r)
    for k in ("bottom", "middle", "top", "quality"):      # same seeds, same order, same quality
        assert np.array_equal(ours[k].view(np.uint32), cpu[k].view(np.uint32)), k
    ref_cuda = oracle.ref_adapter_run(0, ev, resizable_input=resizable)
    a, b = _set(ours), _set(ref_cuda)
    assert len(a & b) >= 0.999 * max(len(a), len(b)), (len(a), len(b), len(a & b))
    # parameters, read through the reference type's accessors
    rp = oracle.ref_estimate_params(cpu["bottom"], cpu["middle"], cpu["top"], ev.xyz, ev.bfield,
                                    sp_meas_index=ev.meas_index, meas_local=ev.meas_local,
                                    meas_surface=ev.meas_surface)
    p = ours["params"]
    assert np.array_equal(p["surface_link"], rp["surface_link"])
    assert np.array_equal(p["vec"][:, :2], rp["vec"][:, :2])
    assert rel_close(p["vec"], rp["vec"]).all()
    diag = np.arange(6) * 7
    assert rel_close(p["cov"][:, diag], rp["cov"][:, diag]).all()
    off = np.ones(36, bool)
    off[diag] = False
    assert not p["cov"][:, off].any()


def test_adapter_empty_event():
    ev = toy_detector.ToyEvent(np.zeros((0, 3), np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32),
                               np.zeros(0, np.uint32), np.zeros((0, 2), np.float32), np.zeros(0, np.uint64),
                               np.zeros(0, np.uint32), 0, np.array([0, 0, 5.9e-4], np.float32))
    ours = oracle.ref_adapter_run(1, ev)
    assert len(ours["bottom"]) == 0
