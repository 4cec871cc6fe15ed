"""Cluster (CSI) volumes through the product's host shim ON THE CUDA ENGINE: the scenarios and random streams of
tests/test_shim_volumes_cpu.py with the shim linked against libplacement.so instead of the oracle ABI.

Written after this round's GPU minutes were spent: the file has not run on a B200 yet.  It is the last file of the suite
and its tests are non-strict xfail, so that whatever they show on the first GPU run is recorded (XPASS = the volume path
works on the device; xfailed = it does not, with the traceback under -rx) without stopping the `-x` run of everything
that HAS been measured.  Drop the marker once a GPU run has passed.  The CPU twin pins the same shim source decision for
decision; what is new on the device are primitives the GPU suite covers elsewhere (one-task and k-task groups restricted
by a leaf term: test_parity_gpu.py::test_random_leaf_visits; batched pe_fit: the preassigned known answers)."""
import pytest

import tests.test_shim_volumes_cpu as cpu_twin
from swarmkit_b200 import _build
from tests.sched_harness import JsonScheduler

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300),
              pytest.mark.xfail(strict=False, reason="first GPU run of the volume path (written with no GPU minutes left); see the module docstring")]


def make_mirror():
    return JsonScheduler(_build.build_scheduler_shim(), "ss")


@pytest.mark.parametrize("name", cpu_twin.SCENARIOS)
def test_reference_volume_scenarios_on_gpu(name, monkeypatch):
    monkeypatch.setattr(cpu_twin, "make_shim", make_mirror)
    cpu_twin.test_reference_volume_scenarios_through_the_shim(name, monkeypatch)


@pytest.mark.parametrize("seed", range(4))
def test_random_event_stream_with_volumes_on_gpu(seed, monkeypatch):
    monkeypatch.setattr(cpu_twin, "make_shim", make_mirror)
    cpu_twin.test_random_event_stream_with_volumes(seed)


@pytest.mark.parametrize("seed", range(4))
def test_random_replicated_services_on_static_volumes_on_gpu(seed, monkeypatch):
    monkeypatch.setattr(cpu_twin, "make_shim", make_mirror)
    cpu_twin.test_random_replicated_services_on_static_volumes(seed)


@pytest.mark.parametrize("seed", range(4))
def test_random_replicated_services_on_volumes_that_count_their_users_on_gpu(seed, monkeypatch):
    monkeypatch.setattr(cpu_twin, "make_shim", make_mirror)
    cpu_twin.test_random_replicated_services_on_volumes_that_count_their_users(seed)


def test_three_writers_one_single_writer_volume_on_gpu(monkeypatch):
    monkeypatch.setattr(cpu_twin, "make_shim", make_mirror)
    cpu_twin.test_three_writers_one_single_writer_volume()


def test_volume_node_set_is_marked_incrementally_on_gpu(monkeypatch):
    monkeypatch.setattr(cpu_twin, "make_shim", make_mirror)
    cpu_twin.test_volume_node_set_is_marked_incrementally()
