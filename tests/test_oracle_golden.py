"""Pins the oracle: the plain-C restatement (oracle/grid_oracle.c) must reproduce
all 13 golden `.task` vectors of the reference at the reference's own unit-test
tolerance (src/grid/grid_unittest.c:51-56: 1e-12), in the same four modes the
reference runs them (collocate/integrate x single-product/batched,
src/grid/grid_unittest.c:80-92)."""
import pytest

from replay import TASK_NAMES, load_task, replay_batched, replay_single_oracle

TOL = 1e-12


def test_all_thirteen_vectors_present():
    assert len(TASK_NAMES) == 13


@pytest.mark.parametrize("name", TASK_NAMES)
@pytest.mark.parametrize("collocate", [True, False], ids=["collocate", "integrate"])
def test_oracle_single_product(oracle, name, collocate):
    assert replay_single_oracle(oracle, load_task(name), collocate) < TOL


@pytest.mark.parametrize("name", TASK_NAMES)
@pytest.mark.parametrize("collocate", [True, False], ids=["collocate", "integrate"])
def test_oracle_batched(oracle, name, collocate):
    assert replay_batched(oracle, load_task(name), collocate) < TOL


@pytest.mark.parametrize("name", ["ortho_density_l0122", "general_tau"])
@pytest.mark.parametrize("collocate", [True, False], ids=["collocate", "integrate"])
def test_oracle_batched_many_cycles(oracle, name, collocate):
    # 7 tasks over 4 blocks; tolerance scales with cycles (grid_miniapp.c:71)
    assert replay_batched(oracle, load_task(name), collocate, cycles=7, cycles_per_block=2) < 7 * TOL


@pytest.mark.parametrize("name", TASK_NAMES)
@pytest.mark.parametrize("collocate", [True, False], ids=["collocate", "integrate"])
def test_reference_ref_backend_batched(reference, name, collocate):
    """The unmodified reference (REF backend) through the same harness: proves
    the harness itself (block layout, scaling, comparison) is faithful."""
    lib = reference.load_reference()
    assert replay_batched(lib, load_task(name), collocate, cycles=3, cycles_per_block=2) < 3 * TOL
