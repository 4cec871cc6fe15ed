"""Pins the object-level CPU oracle (oracle/sched_oracle.cpp) against the
reference's own known-answer tests."""
import inspect

import pytest

from tests import known_answers as KA
from tests.oracle_lib import build_sched
from tests.sched_harness import JsonScheduler


def make():
    return JsonScheduler(build_sched(), "so")


SCENARIOS = [(n, f) for n, f in inspect.getmembers(KA, inspect.isfunction) if n.startswith("scenario_")]


@pytest.mark.parametrize("name,fn", SCENARIOS, ids=[n for n, _ in SCENARIOS])
def test_known_answer(name, fn):
    params = list(inspect.signature(fn).parameters)
    if "use_spec_version" in params:
        for v in (False, True):
            fn(make, v)
    else:
        fn(make)
