// batch.hpp — lane-tiled batch runs and the multi-GPU output gather behind the C ABI (b2_batch_*, b2_comm_*).
//
// A batch of L stimulus lanes is L independent gsim simulators of one netlist (reference: one Simulator::run_sim per
// stimulus vector, crates/digilogic_gsim/src/lib.rs:216-219).  The batch runner owns the tiling of the lanes, the
// simulator replicas that work on alternate tiles, the stimulus source and the destinations of the designated output
// rows — what digilogic_b200/batch.py did in Python + torch in round 1.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "engine.hpp"

namespace b2 {

struct SweepSeg;
struct SweepArgs;

// NCCL communicator of the stimulus-sharded run (one rank per GPU).  libnccl is resolved at run time (dlopen, the copy
// already loaded into the process first), so the library has no link-time dependency on it.
class Comm {
   public:
    static int unique_id(uint8_t id[128]);
    static int create(Comm **out, int n_ranks, int rank, const uint8_t id[128], int device);
    ~Comm();
    int n_ranks() const { return n_ranks_; }
    int rank() const { return rank_; }
    int device() const { return device_; }
    // in-place all-gather: every rank's `count` u64 words at buf + rank * count
    int all_gather_u64(uint64_t *buf, size_t count, void *stream);
    int all_reduce_max_u32(uint32_t *buf, size_t count, void *stream);

   private:
    Comm() = default;
    void *comm_ = nullptr;  // ncclComm_t
    int n_ranks_ = 1, rank_ = 0, device_ = 0;
};

struct BatchInfo {
    uint32_t tile_words, n_slots;
    uint32_t sweep;            // 1: the persistent level-sweep kernel is selected and serves this netlist, 0: simulator replicas
    uint32_t tasks_per_pass;   // sweep: tasks of one tile
    uint32_t grid_ctas, gates_per_task;
    uint32_t n_ranks, rank, peers_mapped;
    uint32_t pair_launches;    // replicas: launches per tile of the level-pair plan (0 = one per level)
    uint64_t store_bytes;
    uint64_t words_local;      // lane-words per rank of the gathered planes (0 without a communicator)
    uint64_t last_tiles, last_tickets;
};

enum BatchOption {
    BATCH_OPT_SWEEP = 1,       // default 0; 1: use the persistent level-sweep kernel when the netlist allows it
    BATCH_OPT_TASK_ITERS = 2,  // default 1: iterations (of 256 threads x U gates) per gate task
    BATCH_OPT_UNROLL = 3,      // default 4: gates per thread of the one/two-input class (2 or 4)
    BATCH_OPT_CTAS_PER_SM = 4, // default 0 = occupancy limit
    BATCH_OPT_AUTO_GATHER = 5, // default 1: with a communicator bound, every run ends with gather_outputs()
    BATCH_OPT_DISCARD = 6      // default 1: replicas drop rows from L2 once their last reader has run (outputs are kept)
};

class Batch {
   public:
    Batch(const Engine &proto, const uint32_t *inputs, uint32_t n_in, const uint32_t *outputs, uint32_t n_out, uint32_t tile_words,
          uint32_t n_slots);
    ~Batch();
    Batch(const Batch &) = delete;
    Batch &operator=(const Batch &) = delete;

    int validate() const;  // wire ids, 1-bit inputs / outputs
    int set_device(int device);
    int set_stream(void *stream);
    int set_option(int option, int64_t value);
    int info(BatchInfo *out);

    // multi-GPU: gathered planes [n_ranks][n_out][words_local] on every rank
    int bind_comm(Comm *comm, uint64_t words_local);
    int export_gather_handle(uint8_t handle[128]);
    int import_gather_handles(const uint8_t *handles);
    int gathered(uint64_t **val, uint64_t **vld);
    int gather_outputs();

    // one settle of lanes [0, 64 * n_words) of the batch; stimulus generated on the device (stim_mode 0 / 1) or read from
    // planes [n_in][in_stride] (device or page-locked host memory; in_ring: words per row when used as a ring)
    int run(int stim_mode, uint64_t seed, uint64_t word_begin, const uint64_t *in_val, const uint64_t *in_vld, uint64_t in_stride,
            uint64_t in_ring, uint64_t n_words, uint64_t max_steps, uint64_t *out_val, uint64_t *out_vld, uint64_t out_stride,
            uint64_t out_ring);
    int wait(int *worst);

   private:
    int ensure_device();
    int ensure_sweep();
    int ensure_replicas();
    int run_sweep(int stim_mode, uint64_t seed, uint64_t word_begin, const uint64_t *in_val, const uint64_t *in_vld, uint64_t in_stride,
                  uint64_t in_ring, uint64_t n_words, uint64_t *out_val, uint64_t *out_vld, uint64_t out_stride, uint64_t out_ring);
    int run_replicas(int stim_mode, uint64_t seed, uint64_t word_begin, const uint64_t *in_val, const uint64_t *in_vld,
                     uint64_t in_stride, uint64_t in_ring, uint64_t n_words, uint64_t max_steps, uint64_t *out_val, uint64_t *out_vld,
                     uint64_t out_stride, uint64_t out_ring);
    int device_pointer(const void *p, size_t bytes, void **dev, bool *is_host);
    void release_registered();

    std::shared_ptr<NetShare> share_;
    const Compiled &net_;
    std::shared_ptr<DeviceNet> dnet_;
    std::unique_ptr<Engine> proto_;  // carries the options of the simulator the batch was made from
    std::vector<uint32_t> inputs_, outputs_;
    uint32_t T_, K_;
    int device_ = -1;
    bool device_ready_ = false;
    void *stream_ = nullptr;
    bool own_stream_ = false, stream_set_ = false;
    int sm_count_ = 148;
    // options
    bool opt_sweep_ = false;
    uint32_t opt_iters_ = 1, opt_unroll_ = 4, opt_ctas_per_sm_ = 0;
    bool opt_auto_gather_ = true, opt_discard_ = true;
    // sweep state
    bool sweep_eligible_ = false, sweep_ready_ = false;
    uint64_t *val_ = nullptr, *vld_ = nullptr;
    SweepSeg *d_segs_ = nullptr;
    uint32_t *d_seg_task0_ = nullptr, *d_in_rows_ = nullptr, *d_out_rows_ = nullptr;
    unsigned long long *d_sync_ = nullptr;
    uint32_t n_segs_ = 0, tasks_per_pass_ = 0, grid_ = 0, gates_per_task_ = 0;
    uint64_t store_bytes_ = 0;
    uint32_t built_iters_ = 0, built_unroll_ = 0;
    // replicas (netlists the sweep does not serve, and max_steps below the depth)
    std::vector<std::unique_ptr<Engine>> replicas_;
    std::vector<void *> rstreams_, revents_;
    void *ev_fork_ = nullptr;
    // results
    int worst_ = 0;
    uint32_t *d_status_ = nullptr, *h_status_ = nullptr;
    uint64_t last_tiles_ = 0, last_tickets_ = 0;
    // multi-GPU
    Comm *comm_ = nullptr;
    uint64_t words_local_ = 0;
    uint64_t *gat_val_ = nullptr, *gat_vld_ = nullptr;
    std::vector<uint64_t *> peer_val_, peer_vld_;  // [n_ranks]; own entry = own buffer
    bool peers_mapped_ = false, gather_pending_ = false;
    std::vector<void *> registered_;  // host ranges page-locked by this batch for the current run
};

}  // namespace b2
