// Example user reward hook (see jsl_engine.cuh "User hook").  Dense shaping: every env.step additionally earns
// -(time advanced by this step) / max_allowed_time, on top of the stock BinaryActionJsspReward.
// The engine hands in the state AFTER the step; the time before it is not kept, so the hook uses the number of
// busy machines as a proxy -- a deliberately simple example of reading per-machine state on the device:
//     reward' = stock_reward - 0.001 * (#machines not WORKING) / n_machines
template <int JW, bool ST, int CL>
JSL_D double jsl_user_reward(const EnvS<JW, ST, CL>& s, const Ctx& c, double stock_reward) {
    const int M = c.I->M;
    const int idle = wp_ballot<1>(M, [&](int m) { return s.m_state[m] != JSL_M_WORKING; }).count();
    return stock_reward - 0.001 * (double)idle / (double)M;
}
