// Grouped launches: ONE launch of a convolution kernel covers the same layer of K experts (north_star: "all K experts are
// batched into one grouped launch per timestep").  gridDim.y indexes the expert; every expert keeps its own parameter block
// and TMA tensor maps (its weights, its workspace), selected by blockIdx.y, and gets gridDim.x persistent CTAs.  At the
// benchmark batch every launch already fills the machine many times over; grouping matters when the per-GPU batch is small
// (strong scaling: 4096 samples over 8 GPUs = 512 each), where it doubles the tiles per launch and halves the launches.
//
// Mechanics: the expert graphs call the ordinary per-expert launchers.  Between group_begin() and group_flush() those
// launchers do not launch; their innermost instance functions RECORD (kernel instance, tensor maps, parameter block, grid,
// shared memory).  group_flush() checks that the K records name the same kernel instance and issues one grouped launch
// (or falls back to K plain launches when they do not).
#pragma once
#include <cuda.h>

#include "cdm_common.cuh"

namespace cdm {

constexpr int GROUP_MAX = 4;
enum GroupKind { GK_NONE = 0, GK_HALO = 1, GK_STACK3 = 2 };

template <int NMAPS, typename P> struct GroupArgs {
  CUtensorMap tm[GROUP_MAX][NMAPS];
  P p[GROUP_MAX];
};

struct GroupRec {
  int kind, inst;
  CUtensorMap tm[6];
  alignas(16) unsigned char params[384];
  int grid;
  size_t smem;
  double flops, bytes;
  char tag[56];
};
struct GroupState {
  bool recording = false;
  int n = 0;
  GroupRec rec[GROUP_MAX];
};
inline GroupState& group_state() {
  static thread_local GroupState s;
  return s;
}
inline bool group_recording() { return group_state().recording; }
inline void group_begin() { GroupState& g = group_state(); g.recording = true; g.n = 0; }
// record one launch; returns false when the group is full (the caller then launches directly)
template <typename P> inline GroupRec* group_record(int kind, int inst, const P& p, int grid, size_t smem, double flops, double bytes,
                                                    const char* tag) {
  static_assert(sizeof(P) <= sizeof(GroupRec::params), "parameter block too large for a group record");
  GroupState& g = group_state();
  if (g.n >= GROUP_MAX) return nullptr;
  GroupRec& r = g.rec[g.n++];
  r.kind = kind; r.inst = inst; r.grid = grid; r.smem = smem; r.flops = flops; r.bytes = bytes;
  memcpy(r.params, &p, sizeof(P));
  r.tag[0] = 0;
  if (tag) { strncpy(r.tag, tag, sizeof(r.tag) - 1); r.tag[sizeof(r.tag) - 1] = 0; }
  return &r;
}
// kernels that can run grouped (defined next to their kernels)
int launch_halo_group(const GroupRec* recs, int K, int num_sms, cudaStream_t st);
int launch_stack3_group(const GroupRec* recs, int K, int num_sms, cudaStream_t st);
// ends the recording and launches what was recorded (one grouped launch when the K records agree)
int group_flush(int num_sms, cudaStream_t st);

}  // namespace cdm
