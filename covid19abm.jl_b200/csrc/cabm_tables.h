// cabm_tables.h — host-side draw-contract tables (see cabm_tables.cpp)
#pragma once
#include <cstdint>
#include <vector>

namespace cabm {

struct Table {
    int base = 0;
    std::vector<uint32_t> thr;   // value = base + #{k : thr[k] <= w}
};
struct Tables {
    Table lat, pre, asy, inf, psi, mu, nb[5];
};

const Tables &tables();
const Table *table_by_name(const char *name);
void gpw_split(int ag, int cnts, int out[5]);       // src/covid19abm.jl:804, ag 1..5
uint64_t bernoulli_threshold(double p);             // rand() < p  <=>  w < threshold
extern const double CONTACT_MATRIX[5][5];           // src/covid19abm.jl:838-848
extern const double PROV_AG[15][5];                 // src/covid19abm.jl:319-333

}  // namespace cabm
