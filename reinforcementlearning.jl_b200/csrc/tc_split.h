// tc_split.h — how the tensor-core loss + backward kernel (nn_tc.cu) divides its persistent CTAs between the actor and the critic.
// Plain C++ (also compiled for the host by the CPU test suite).
#pragma once
#include <cstdint>

// CTAs given to the actor out of `grid` (the rest work on the critic).  A critic tile costs ~0.87 of an actor tile with the
// categorical PPO / A2C loss and ~0.85 with the Gaussian head (B200 sweeps, profiles/na_sweep.sh: 79 : 69 is the optimum for the
// 4 096-tile BASELINE minibatch, 80 : 68 for the 8 192-tile Pendulum batch); the split minimises the longer of the two roles'
// whole-tile counts.  At most grid / 2 + 8 (the fused optimiser step stages <= 82 partial rows), at least grid / 2.
static inline int b200rl_tc_actor_ctas(int grid, bool gaussian_head, int64_t ntiles) {
    const double r = gaussian_head ? 0.85 : 0.87;
    int best = grid / 2;
    double best_cost = 1e300;
    for (int na = grid / 2; na <= grid / 2 + 8 && na < grid; ++na) {
        const double ca = (double)((ntiles + na - 1) / na), cc = r * (double)((ntiles + (grid - na) - 1) / (grid - na));
        const double cost = ca > cc ? ca : cc;
        if (cost < best_cost - 1e-9) { best_cost = cost; best = na; }
    }
    return best;
}
