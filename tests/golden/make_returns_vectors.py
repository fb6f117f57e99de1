"""Writes tests/golden/returns_reference_vectors.json: the reference's OWN known-answer vectors for
`discount_rewards`, `discount_rewards_reduced` and `generalized_advantage_estimation`, transcribed from the
assertions of /root/reference/src/ReinforcementLearningCore/test/utils/base.jl (lines 22-152).  The reference is
Julia and cannot run in the build image, so nothing is *computed* here: each entry is the literal input /
expected output of one `@test` there (`ref` = line), matrices in Julia's row-by-row literal order.

    python tests/golden/make_returns_vectors.py          # rewrites the JSON next to this file
"""
import json
import os

R9 = {"reshape_1_to": 9, "shape": [3, 3]}      # reshape(1:9, 3, 3), column-major
V43 = {"reshape_1_to": 12, "shape": [4, 3]}    # reshape(1:12, 4, 3)
V34 = {"reshape_1_to": 12, "shape": [3, 4]}    # reshape(1:12, 3, 4)
TERM = [[0, 1, 0], [1, 0, 1], [0, 1, 0]]
INIT = [-2.0, 0.0, 2.0]

discount = [  # base.jl:22-62
    dict(ref=24, rewards=[1.0], gamma=0.5, expected=[1.0]),
    dict(ref=25, rewards=[1.0], gamma=0.5, init=2.0, expected=[2.0]),
    dict(ref=29, rewards=[1, 2, 3], gamma=0.5, expected=[2.75, 3.5, 3.0]),
    dict(ref=30, rewards=[1, 2, 3], gamma=0.5, init=4.0, expected=[3.25, 4.5, 5.0]),
    dict(ref=32, rewards=[1, 2, 3], gamma=0.5, terminal=[0, 1, 0], init=2.0, expected=[2.0, 2.0, 4.0]),
    dict(ref=34, rewards=[1, 2, 3], gamma=0.5, terminal=[1, 0, 1], init=2.0, expected=[1.0, 3.5, 3.0]),
    dict(ref=46, rewards=R9, gamma=0.5, dims=1, expected=[[2.75, 8.0, 13.25], [3.5, 8.0, 12.5], [3.0, 6.0, 9.0]]),
    dict(ref=48, rewards=R9, gamma=0.5, dims=2, expected=[[4.75, 7.5, 7.0], [6.5, 9.0, 8.0], [8.25, 10.5, 9.0]]),
    dict(ref=50, rewards=R9, gamma=0.5, init=INIT, dims=1, expected=[[2.5, 8.0, 13.5], [3.0, 8.0, 13.0], [2.0, 6.0, 10.0]]),
    dict(ref=52, rewards=R9, gamma=0.5, init=INIT, dims=2, expected=[[4.5, 7.0, 6.0], [6.5, 9.0, 8.0], [8.5, 11.0, 10]]),
    dict(ref=56, rewards=R9, gamma=0.5, dims=1, terminal=TERM, expected=[[2.0, 4.0, 11.0], [2.0, 8.0, 8.0], [3.0, 6.0, 9.0]]),
    dict(ref=58, rewards=R9, gamma=0.5, dims=1, terminal=TERM, init=INIT, expected=[[2.0, 4.0, 11.0], [2.0, 8.0, 8.0], [2.0, 6.0, 10.0]]),
    dict(ref=60, rewards=R9, gamma=0.5, dims=2, terminal=TERM, init=INIT, expected=[[3.0, 4.0, 6.0], [2.0, 9.0, 8.0], [6.0, 6.0, 10.0]]),
]
reduced = [  # base.jl:64-102
    dict(ref=66, rewards=[1.0], gamma=0.5, expected=1.0),
    dict(ref=69, rewards=[1, 2, 3], gamma=0.5, expected=2.75),
    dict(ref=70, rewards=[1, 2, 3], gamma=0.5, init=4.0, expected=3.25),
    dict(ref=71, rewards=[1, 2, 3], gamma=0.5, terminal=[0, 1, 0], expected=2.0),
    dict(ref=72, rewards=[1, 2, 3], gamma=0.5, terminal=[0, 1, 0], init=4.0, expected=2.0),
    dict(ref=86, rewards=R9, gamma=0.5, dims=1, expected=[2.75, 8.0, 13.25]),
    dict(ref=87, rewards=R9, gamma=0.5, dims=2, expected=[4.75, 6.5, 8.25]),
    dict(ref=88, rewards=R9, gamma=0.5, dims=1, terminal=TERM, init=INIT, expected=[2.0, 4.0, 11.0]),
    dict(ref=95, rewards=R9, gamma=0.5, dims=2, terminal=TERM, init=INIT, expected=[3.0, 2.0, 6.0]),
]
gae = [  # base.jl:104-152
    dict(ref=106, rewards=[1.0], values=[2.0, 3.0], gamma=0.5, **{"lambda": 0.3}, expected=[0.5]),
    dict(ref=109, rewards=[1.0, 1.0], values=[1, 2, 3], gamma=0.5, **{"lambda": 0.3}, expected=[1.075, 0.5]),
    dict(ref=112, rewards=[1, 2, 3], values=[1, 2, 3, 4], gamma=0.5, **{"lambda": 0.3}, expected=[1.27, 1.8, 2]),
    dict(ref=114, rewards=[1, 2, 3], values=[1, 2, 3, 4], gamma=0.5, **{"lambda": 0.3}, terminal=[1, 0, 1], expected=[0.0, 1.5, 0.0]),
    dict(ref=130, rewards=R9, values=V43, gamma=0.5, **{"lambda": 0.3}, dims=1, expected=[[1.27, 2.4425, 3.615], [1.8, 2.95, 4.1], [2.0, 3.0, 4.0]]),
    dict(ref=134, rewards=R9, values=V34, gamma=0.5, **{"lambda": 0.3}, dims=2, expected=[[2.6375, 4.25, 5.0], [3.22375, 4.825, 5.5], [3.81, 5.4, 6.0]]),
    dict(ref=140, rewards=R9, values=V43, gamma=0.5, **{"lambda": 0.3}, dims=1, terminal=TERM, expected=[[1.0, -1.0, 2.7], [0.0, 2.35, -2.0], [2.0, -1.0, 4.0]]),
    dict(ref=150, rewards=R9, values=V34, gamma=0.5, **{"lambda": 0.3}, dims=2, expected=[[2.6375, 4.25, 5.0], [3.22375, 4.825, 5.5], [3.81, 5.4, 6.0]]),
]

# Source-derived known answers of the env dynamics (from the formulas at CartPoleEnv.jl:118-140; the
# reference has no golden trajectories): state after 1 and 3 pushes to the right from the zero state.
env_kat = dict(
    cartpole_f64=dict(actions=[2, 2, 2], after_1=[0.0, 0.1951219512195122, 0.0, -0.2926829268292683],
                      after_3=[0.011707317073170733, 0.585447355516259, -0.0175609756097561, -0.8798869825388553]),
    cartpole_f32_bits=dict(actions=[2, 2, 2], after_1=[0x00000000, 0x3E47CE0C, 0x00000000, 0xBE95DA89],
                           after_3=[0x3C3FD00B, 0x3F15DFE0, 0xBC8FDC08, 0xBF614045]),
)

if __name__ == "__main__":
    out = dict(source="ReinforcementLearningCore/test/utils/base.jl:22-152 (reference @ /root/reference)", rtol="sqrt(eps(Float64)) = 1.5e-8 (Julia isapprox default)",
               discount_rewards=discount, discount_rewards_reduced=reduced, generalized_advantage_estimation=gae, env_known_answers=env_kat)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "returns_reference_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path, len(discount), len(reduced), len(gae))
