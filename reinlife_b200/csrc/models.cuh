// models.cuh -- compile-time shapes + flat parameter layout of the three network families
// (see include/reinlife_b200.h "Brains" for the layout contract).
#pragma once
#include "rl_common.cuh"

constexpr int RL_K1 = 160;   // observation row length in floats (153 padded to 160)

template <int KIND> struct Model;
template <> struct Model<RL_MODEL_DUELING> { static constexpr int N1 = 128, N2 = 256, NH = 9; };   // Models/PERD3QN.py:185-202
template <> struct Model<RL_MODEL_DQN>     { static constexpr int N1 = 128, N2 = 64,  NH = 8; };   // Models/DQN.py:119-130
template <> struct Model<RL_MODEL_PPO>     { static constexpr int N1 = 256, N2 = 256, NH = 9; };   // Models/PPO.py:96-112

template <int KIND> struct Layout {
    using M = Model<KIND>;
    static constexpr int OFF_W1T = 0;
    static constexpr int OFF_B1 = RL_K1 * M::N1;
    static constexpr int OFF_W2T = OFF_B1 + M::N1;
    static constexpr int OFF_B2 = OFF_W2T + M::N1 * M::N2;
    static constexpr int OFF_WH = OFF_B2 + M::N2;
    static constexpr int OFF_BH = OFF_WH + M::N2 * M::NH;
    static constexpr int N_TRAIN = (OFF_BH + M::NH + 3) & ~3;
    static constexpr int OFF_W2 = N_TRAIN;                       // output-major copy of W2: [N2][N1]
    static constexpr int N_TOTAL = OFF_W2 + M::N2 * M::N1;
};
