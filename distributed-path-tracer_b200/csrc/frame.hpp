// frame.hpp — entry points of frame.cu (multi-GPU frame driver) used by the C ABI layer (api.cu).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "ptb.h"

struct ptb_scene;
struct ptb_group;
struct ptb_ctx;

namespace ptb {

// one process per GPU: ranks rendezvous through the POSIX shared-memory object "/ptb_<name>"
ptb_group* group_create_processes(const char* name, int rank, int world, int device);
// threads of one process: ranks share a heap block (group_shared_alloc)
void* group_shared_alloc();
void group_shared_free(void* p);
void group_shared_clear_failure(void* p);
ptb_group* group_create_threads(void* shared_block, int rank, int world, int device);
void group_destroy(ptb_group* g);
void group_barrier_public(ptb_group* g);
void group_selftest_host(const char* name, int rank, int world, uint32_t n_tiles, uint32_t frames, uint32_t work_us,
                         uint8_t* mine_out);
// collective: renders the tiles this rank claims straight into rank 0's frame; rank 0 copies the frame out
void group_render_frame(ptb_group* g, const ptb_scene* scene, const ptb_frame_req& req, void* out_host,
                        ptb_frame_stats* stats);
// the last rendered frame in device memory as this rank sees it (tests)
const float4* group_frame_dev(const ptb_group* g);

void frame_tiles(const ptb_frame_req& req, int world, uint32_t* out, uint64_t capacity, uint32_t* n_tiles, bool with_comb);

ptb_ctx* ctx_create(int n_gpus, const int* devices);
void ctx_destroy(ptb_ctx* c);
void ctx_set_scene(ptb_ctx* c, const ptb_scene_desc& desc);
const ptb_scene* ctx_scene(const ptb_ctx* c, int i);
int ctx_size(const ptb_ctx* c);
void ctx_render_frame(ptb_ctx* c, const ptb_frame_req& req, void* out_host, ptb_frame_stats* stats);

} // namespace ptb
