// jsl_tobj.hpp -- host-side flattening of an instance's time objects (DeterministicTimeConfig |
// StochasticTimeConfig, jobshoplab/types/stochasticy_models.py) into one table:
//   ops [0,N) | setup [M*nt*nt] | travel [(M+S)^2] | machine-outage frequency, duration | transport-outage
//   frequency, duration.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/jobshoplab_b200.h"

namespace jsl {

struct TobjTables {
    std::vector<uint8_t> kind;
    std::vector<int32_t> base, init;
    std::vector<float> param;
    int n = 0, off_setup = 0, off_travel = 0, off_mof = 0, off_mod = 0, off_tof = 0, off_tod = 0;
    bool stochastic = false;
    std::vector<int32_t> sample_table, sample_row;  // [rows][depth], [n]
    int sample_depth = 0;
};

inline void tobj_append(TobjTables& t, const int32_t* time, const jsl_time_spec& sp, int n) {
    for (int i = 0; i < n; i++) {
        int k = sp.kind ? sp.kind[i] : JSL_TIME_STATIC;
        t.kind.push_back((uint8_t)k);
        t.init.push_back(time ? time[i] : 0);
        t.base.push_back(sp.base ? sp.base[i] : (time ? time[i] : 0));
        t.param.push_back(sp.param ? sp.param[i] : 0.f);
        if (k != JSL_TIME_STATIC) t.stochastic = true;
    }
    t.n += n;
}

inline TobjTables make_tobj(const jsl_instance_desc& d) {
    TobjTables t;
    const int NP = d.n_machines + d.n_sbuf;
    const int n_mo = d.m_outage_start ? d.m_outage_start[d.n_machines] : 0;
    tobj_append(t, d.op_duration, d.op_spec, d.n_ops);
    t.off_setup = t.n; tobj_append(t, d.setup_time, d.setup_spec, d.n_machines * d.n_tools * d.n_tools);
    t.off_travel = t.n; tobj_append(t, d.travel_time, d.travel_spec, NP * NP);
    t.off_mof = t.n; tobj_append(t, d.m_outage_freq, d.m_ofreq_spec, n_mo);
    t.off_mod = t.n; tobj_append(t, d.m_outage_dur, d.m_odur_spec, n_mo);
    t.off_tof = t.n; tobj_append(t, d.t_outage_freq, d.t_ofreq_spec, d.n_t_outages);
    t.off_tod = t.n; tobj_append(t, d.t_outage_dur, d.t_odur_spec, d.n_t_outages);
    if (t.stochastic && d.sample_table && d.sample_row && d.sample_depth > 0) {
        t.sample_depth = d.sample_depth;
        t.sample_table.assign(d.sample_table, d.sample_table + (size_t)d.n_sample_rows * d.sample_depth);
        t.sample_row.assign(d.sample_row, d.sample_row + t.n);
    }
    return t;
}

}  // namespace jsl
