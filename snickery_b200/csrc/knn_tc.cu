// placeholder until the tcgen05 kernel lands
#include "common.cuh"
bool snk_tc_supported(const snk_db *, const snk_space &, int) { return false; }
int snk_tc_prepare(snk_db *) { return 0; }
void snk_tc_destroy(snk_db *) {}
int snk_tc_query_ld(const snk_db *, int) { return 0; }
const short *snk_tc_qmap(const snk_db *, int) { return nullptr; }
int snk_shortlist_tc(snk_db *, int, const __half *, int, int64_t, int, int, float *, int *, float *, cudaStream_t) {
    snk_set_error("tensor-core engine not built");
    return 1;
}
