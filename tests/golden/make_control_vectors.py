"""Writes tests/golden/control_reference_vectors.json: the reference's OWN known-answer assertions for the host
control plane on the hot path — explorer schedules / probabilities, stop-condition counts, log-prob identities —
transcribed from /root/reference/src/ReinforcementLearningCore/test (`ref` = file:line of the `@test`).  The
reference is Julia and cannot run in the build image, so nothing is *computed* here.

    python tests/golden/make_control_vectors.py          # rewrites the JSON next to this file
"""
import json
import os

EX = dict(eps_init=0.9, eps_stable=0.1, warmup_steps=100, decay_steps=100)
F = "policies/explorers/epsilon_greedy_explorer.jl"
explorer_schedule = [  # get_ϵ(EpsilonGreedyExplorer(kind, ϵ_init = 0.9, ϵ_stable = 0.1, warmup_steps = 100, decay_steps = 100), step)
    dict(ref=f"{F}:8", kind="linear", step=50, expected=0.9),
    dict(ref=f"{F}:9", kind="linear", step=100, expected=0.9),
    dict(ref=f"{F}:10", kind="linear", step=150, expected=0.5),
    dict(ref=f"{F}:11", kind="linear", step=200, expected=0.1),
    dict(ref=f"{F}:15", kind="exp", step=50, expected=0.9),
    dict(ref=f"{F}:17", kind="exp", step=150, expected=0.5852245277701068),
    dict(ref=f"{F}:18", kind="exp", step=2000, expected=0.1, atol=1e-2),
]
VALUES = [0.1, 0.5, 0.5, 0.3]
explorer_prob = [  # prob(s, values[, action]) at step 1 (inside the warm-up: ϵ = 0.9)
    dict(ref=f"{F}:48", is_break_tie=True, values=VALUES, expected=[0.225, 0.275, 0.275, 0.225]),
    dict(ref=f"{F}:49", is_break_tie=True, values=VALUES, action=2, expected=0.275),
    dict(ref=f"{F}:55", is_break_tie=False, values=VALUES, expected=[0.225, 0.32499999999999996, 0.225, 0.225]),
    dict(ref=f"{F}:56", is_break_tie=False, values=VALUES, action=2, expected=0.32500000000000007),
]
greedy = [
    dict(ref=f"{F}:64", values=VALUES, plan=2),
    dict(ref=f"{F}:70", values=VALUES, prob=[0.0, 1.0, 0.0, 0.0]),
    dict(ref=f"{F}:71", values=VALUES, action=2, prob=1.0),
]
explorer_plan_coverage = [  # 300 plan! calls at ϵ = 0.9 visit all four actions
    dict(ref=f"{F}:30", is_break_tie=True, values=VALUES, calls=300, unique_actions=4),
    dict(ref=f"{F}:42", is_break_tie=False, values=VALUES, calls=300, unique_actions=4),
]
S = "core/stop_conditions.jl"
stop_conditions = [  # number of `true` results in `calls` consecutive check! calls
    dict(ref=f"{S}:8", condition=["StopAfterNSteps", 10], calls=20, trues=11),
    dict(ref=f"{S}:22", condition=["StopIfAny", ["StopAfterNSteps", 10], ["StopAfterNSteps", 3]], calls=20, trues=18),
    dict(ref=f"{S}:33", condition=["StopIfAll", ["StopAfterNSteps", 10], ["StopAfterNSteps", 3]], calls=20, trues=11),
    dict(ref=f"{S}:44-47", condition=["StopAfterNEpisodes", 2], sequence=[False, False, True],
         note="one check on a running episode, two on a terminated env"),
    dict(ref=f"{S}:60-62", condition=["StopAfterNoImprovement", "constant 1.0", 10], calls_not_terminated=11, trues_not_terminated=0,
         calls_terminated=11, trues_terminated=1),
]
D = "utils/distributions.jl"
logpdf_identities = [  # `≈ Distributions.logpdf`: the closed forms, evaluated by the test with scipy
    dict(ref=f"{D}:25", fn="normlogpdf", mu=10.0, sigma=5.0, x=4.0, equals="logpdf(Normal(10, 5), 4)"),
    dict(ref=f"{D}:55", fn="diagnormlogpdf", mu=[10.0, 1.0], sigma=[5.0, 6.0], x=[4.0, 3.0],
         equals="logpdf(MvNormal([10, 1], Diagonal([25, 36])), [4, 3])"),
]

out = dict(source="ReinforcementLearningCore/test/{policies/explorers/epsilon_greedy_explorer.jl, core/stop_conditions.jl, utils/distributions.jl} "
                  "(reference @ /root/reference)",
           explorer=EX, explorer_schedule=explorer_schedule, explorer_prob=explorer_prob, greedy=greedy,
           explorer_plan_coverage=explorer_plan_coverage, stop_conditions=stop_conditions, logpdf_identities=logpdf_identities)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "control_reference_vectors.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(path)
