"""Ring-buffer semantics of the trajectory restatement against a plain Python model (hypothesis): whatever the capacity, lane count and
number of pushes, every sampled tuple is a real stored transition — (state, action, reward, terminal) of frame f and the state of frame
f + 1 as next_state (MultiplexTraces, docs/src/How_to_implement_a_new_algorithm.md:90-98) — from the last `capacity` frames only, and the
length follows RLCore/test/policies/agent.jl:27-34.  The GPU ring is then compared with this oracle bit for bit (tests/test_traj_dqn_gpu.py)."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle_lib as O


def obs_of(frame, lanes, ns):
    """state id: [frame * 1000 + lane, -frame, lane, 0.5] — unique per (frame, lane)"""
    o = np.zeros((ns, lanes), np.float32)
    o[0] = frame * 1000 + np.arange(lanes)
    if ns > 1:
        o[1] = -frame
    if ns > 2:
        o[2] = np.arange(lanes)
    return o


@settings(max_examples=40, deadline=None)
@given(lanes=st.integers(1, 9), cap=st.integers(1, 12), pushes=st.integers(1, 40), ns=st.integers(1, 4), prioritized=st.booleans(),
       seed=st.integers(0, 2 ** 20))
def test_every_sample_is_a_stored_transition_of_the_last_capacity_frames(lanes, cap, pushes, ns, prioritized, seed):
    tr = O.OracleTraj(ns, lanes, cap, prioritized, 1.0)
    rng = np.random.default_rng(seed)
    tr.push_state(obs_of(0, lanes, ns))
    assert len(tr) == 0                                             # after the first state: nothing to sample (agent.jl:27-34)
    model = {}
    for f in range(pushes):
        a = rng.integers(1, 4, lanes).astype(np.int32); r = rng.standard_normal(lanes).astype(np.float32)
        t = (rng.random(lanes) < 0.2).astype(np.uint8)
        tr.push(a, r, t, obs_of(f + 1, lanes, ns))
        model[f] = (a, r, t)
        assert len(tr) == min(f + 1, cap)
    B = 64
    slots = O.splitmix_states_fast(B, seed)
    if prioritized:
        assert tr.total_priority() == float(min(pushes, cap) * lanes)      # default priority 1 on every sampleable transition, 0 elsewhere
    b = tr.sample(slots, B, prioritized=prioritized, beta=0.5)
    oldest = pushes - min(pushes, cap)
    for k in range(B):
        sid = int(b["state"][0, k])
        f, e = divmod(sid, 1000)
        assert oldest <= f < pushes and 0 <= e < lanes
        a, r, t = model[f]
        assert b["action"][k] == a[e] and b["reward"][k] == r[e] and b["terminal"][k] == t[e]
        assert np.array_equal(b["state"][:, k], obs_of(f, lanes, ns)[:, e])
        assert np.array_equal(b["next_state"][:, k], obs_of(f + 1, lanes, ns)[:, e])        # :next_state at i is :state at i + 1
    assert len(np.unique(b["key"])) == len({(int(b["state"][0, k])) for k in range(B)})      # key <-> transition is one to one
    if prioritized:
        assert np.all(b["priority"] == 1.0) and np.allclose(b["weight"], 1.0)
        # zeroing the priority of everything sampled makes those transitions unreachable
        tr.update_priority(b["key"], np.zeros(B, np.float32))
        if tr.total_priority() > 0:
            b2 = tr.sample(slots, B, prioritized=True, beta=0.5)
            assert not set(b2["key"].tolist()) & set(b["key"].tolist())
