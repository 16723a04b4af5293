"""Host-side logic of gridcodegenerator_b200/runtime.py that needs no GPU: argument validation of the
device wrappers and the environment -> grid_set_option synchronisation."""
import os

import pytest

torch = pytest.importorskip("torch")

from gridcodegenerator_b200 import load_named_robot                      # noqa: E402
from gridcodegenerator_b200.runtime import GridEngine, GridError        # noqa: E402


@pytest.fixture(scope="module")
def eng():
    return GridEngine(load_named_robot("iiwa14"))


class FakeCuda:
    """Stands in for a CUDA tensor on a box without a GPU: only what runtime._shape/_ptr read."""

    def __init__(self, *shape, device="cuda:0"):
        self.shape, self._n = shape, 1
        for s in shape:
            self._n *= s
        self.device = torch.device(device)
        self.is_cuda, self.dtype = True, torch.float32

    def dim(self):
        return len(self.shape)

    def numel(self):
        return self._n

    def is_contiguous(self):
        return True

    def data_ptr(self):
        return 1 << 20


def test_shape_validation_rejects_mismatches(eng, monkeypatch):
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    n = eng.n
    x = FakeCuda(16, 3 * n)
    ok = eng._shape("t", x, 3 * n, None, None, [("out", FakeCuda(16, 2 * n * n), 2 * n * n)])
    assert ok == (16, 3 * n)
    with pytest.raises(GridError, match="2-D"):
        eng._shape("t", FakeCuda(3 * n), 3 * n, None, None, [])                         # 1-D single state
    with pytest.raises(GridError, match="holds"):
        eng._shape("t", x, 3 * n, None, None, [("out", FakeCuda(15, 2 * n * n), 2 * n * n)])    # output too small
    with pytest.raises(GridError, match="stride"):
        eng._shape("t", FakeCuda(16, 2 * n), 3 * n, None, None, [])                     # rows shorter than 3n
    with pytest.raises(GridError, match="fewer"):
        eng._shape("t", x, 3 * n, 17, 3 * n, [])                                        # more states than rows
    with pytest.raises(GridError, match="is on"):
        eng._shape("t", x, 3 * n, None, None, [("out", FakeCuda(16, 2 * n * n, device="cuda:1"), 2 * n * n)])
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 1)
    with pytest.raises(GridError, match="current device"):
        eng._shape("t", x, 3 * n, None, None, [])
    with pytest.raises(GridError, match="num_timesteps"):
        eng._shape("t", 1 << 20, 3 * n, None, 3 * n, [])                                # raw pointer without a count


def test_environment_changes_reach_the_library(eng, monkeypatch):
    monkeypatch.setenv("GRID_FORCE_KERNEL", "cps")
    eng._sync_options()
    assert eng._pushed_options[0] == "cps"
    monkeypatch.setenv("GRID_FORCE_KERNEL", "not-a-family")
    with pytest.raises(GridError, match="GRID_FORCE_KERNEL"):
        eng._sync_options()
    monkeypatch.delenv("GRID_FORCE_KERNEL")
    eng._sync_options()
    assert eng._pushed_options[0] is None
    eng.set_option("GRID_PIPE_CHUNK", "4096")
    assert os.environ["GRID_PIPE_CHUNK"] == "4096"
    eng.set_option("GRID_PIPE_CHUNK", None)
    assert "GRID_PIPE_CHUNK" not in os.environ


def test_host_tensors_are_rejected(eng):
    n = eng.n
    with pytest.raises(GridError):
        eng.forward_dynamics_gradient_device(torch.empty(4, 2 * n * n), torch.empty(4, 3 * n))
