// stubs.cu -- entry points still under construction; they fail loudly rather than fall back.
#include <vector>
#include "engine.hpp"
#include "stages.hpp"
using namespace ckm;
extern "C" {
int ckm_allgather_qa(ckm_engine *, void *, const ckm_qa_row *, int32_t, int32_t, int32_t, ckm_qa_row *, int32_t *) { set_error("ckm_allgather_qa: not implemented yet"); return CKM_EINVAL; }
}
extern "C" {
}

