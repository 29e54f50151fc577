// migrate.cu -- multi-GPU slab exchange.  Filled in below.
#include "common.cuh"

using namespace fsim;

extern "C" {

int64_t fsim_migrate_record_bytes(const fsim_sim *s) { return s ? (int64_t)(NPART_ARRAYS * s->rs + 8) : -1; }

int fsim_migrate_pack(fsim_sim *, const int64_t *, int32_t, int32_t, int64_t *, void **)
{
    set_error("fsim_migrate_pack: not implemented yet");
    return FSIM_ERR_UNSUPPORTED;
}
int fsim_migrate_unpack(fsim_sim *, const void *, int64_t)
{
    set_error("fsim_migrate_unpack: not implemented yet");
    return FSIM_ERR_UNSUPPORTED;
}
int fsim_halo_ptrs(fsim_sim *, void **, void **, void **, void **, int64_t *)
{
    set_error("fsim_halo_ptrs: not implemented yet");
    return FSIM_ERR_UNSUPPORTED;
}

}  // extern "C"
