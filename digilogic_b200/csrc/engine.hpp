// engine.hpp — device-resident simulator: wire-state store, drive store, settle loops.
//
// Replaces gsim::Simulator behind the calls the reference makes at
// crates/digilogic_gsim/src/lib.rs:209-230 (set_wire_drive, run_sim, get_wire_state).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "netlist.hpp"

namespace b2 {

void set_last_error(const std::string &msg);
const char *get_last_error();
uint64_t kernel_launches();

struct RunFlags {  // written by the status kernel, read by the host after each application
    uint32_t any_live, any_conflict, any_unsettled;
    uint32_t active;  // gates listed for this application (frontier lists); 0xFFFFFFFF: every gate was evaluated
};

struct LoopResult;

// The compiled netlist on one device: index and kind arrays the kernels stream.  Read-only after the upload
// and shared by every simulator replica built from one b2_build (b2_sim_clone, b2_batch).
struct DeviceNet {
    int device = -1;
    uint32_t *fan0 = nullptr, *fan1 = nullptr, *fan2 = nullptr, *wide_off = nullptr, *wide_in = nullptr;
    uint8_t *kind2 = nullptr, *far2 = nullptr, *wide_kind = nullptr, *wide_nsel = nullptr, *wide_aux = nullptr;
    uint32_t *res_off = nullptr, *res_in = nullptr;
    Level *levels = nullptr;
    ~DeviceNet();
};

// Level-pair fusion plan of a netlist on one device (pair.cuh): items, discard lists, launch table.  Shared by the
// replicas of a batch (same netlist, same kept rows).
struct PairLaunch;
struct PairPlanDev {
    int device = -1;
    std::vector<uint32_t> keep_rows;  // sorted: the rows the plan keeps (its key)
    bool keep_all = false;            // ... or every row (plain simulator runs)
    void *items = nullptr;            // PairItem[] on the device
    uint32_t *disc_rows = nullptr;
    std::vector<uint32_t> launch_tab;  // 4 per launch: item_begin, item_end, disc_begin, disc_end
    uint64_t n_items = 0, n_children = 0, n_unstored = 0;
    ~PairPlanDev();
};

// One compiled netlist (host arrays) and its per-device copies.
struct NetShare {
    explicit NetShare(Compiled &&c) : net(std::move(c)) {}
    const Compiled net;
    std::mutex m;
    std::map<int, std::weak_ptr<DeviceNet>> on_device;
    // the device copy for `device` (current CUDA device must be `device`); uploaded on first use; rc = b2_status
    std::shared_ptr<DeviceNet> acquire(int device, int &rc);
    std::map<int, std::weak_ptr<PairPlanDev>> pair_plans;  // per device: the last plan built (reused when the kept rows match)
};

class Engine {
   public:
    explicit Engine(Compiled &&net);
    explicit Engine(std::shared_ptr<NetShare> share);  // a replica: shares the netlist and its device copy
    ~Engine();
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;

    const Compiled &net() const { return net_; }
    const std::shared_ptr<NetShare> &share() const { return share_; }
    int device() const { return device_; }
    // engine options as set so far (copied into replicas)
    void copy_options_from(const Engine &o);

    int set_device(int device);
    int set_stream(void *stream);
    int synchronize();
    int set_lanes(uint64_t n_lanes);
    uint64_t max_lanes(uint64_t bytes) const;

    int set_wire_drive(uint32_t wire, const uint32_t *p0, const uint32_t *p1, size_t words);
    int set_wire_drive_lanes(uint32_t wire, uint32_t bit, uint32_t w0, uint32_t nw, const uint64_t *value, const uint64_t *valid);
    int set_drives_lanes(const uint32_t *wires, uint32_t n, const uint64_t *value, const uint64_t *valid, size_t src_stride);
    int fill_stimulus(const uint32_t *wires, uint32_t n, uint64_t seed, uint64_t word_base, int mode);

    // returns b2_run_result or -(b2_status)
    int run(uint64_t max_steps, uint64_t *conflict, uint64_t *unsettled, bool async);

    int get_wire_state(uint32_t wire, uint32_t *p0, uint32_t *p1, size_t words);
    int get_wire_lanes(uint32_t wire, uint32_t bit, uint32_t w0, uint32_t nw, uint64_t *value, uint64_t *valid);
    int gather_wires(const uint32_t *wires, uint32_t n, uint64_t *value, uint64_t *valid, size_t dst_stride, bool dst_is_device,
                     bool async = false);
    // Output rows of bit 0 of n wires to several DEVICE destinations at once (own planes, peer GPUs' planes) with one kernel on
    // a side stream: the rows are read once, and the simulator's stream goes on with its next run meanwhile (the next run's
    // first write to the store waits for the side stream).
    struct GatherDst {
        uint64_t *val, *vld;
        uint64_t stride;
    };
    int gather_wires_multi(const uint32_t *wires, uint32_t n, const GatherDst *dsts, uint32_t n_dst);
    int join_copies();
    // every lane back to the state of a freshly built simulator (wires and component outputs as at build time); drives are kept
    void reset_cold() { cold_ = true; frontier_valid_ = false; snap_valid_ = false; }
    int pack_sim_state(const uint32_t *wires, uint32_t n, uint8_t *p0, uint8_t *p1, size_t cap, uint64_t *bit_len);

    // info
    uint64_t n_lanes() const { return n_lanes_; }
    uint32_t lane_words() const { return T_; }
    uint32_t row_stride() const { return stride_; }
    uint64_t store_bytes() const { return store_bytes_; }
    uint32_t last_mode() const { return last_mode_; }
    uint64_t last_applications() const { return last_apps_; }
    void set_use_resident(bool on) { use_resident_ = on; }
    void set_valid_elision(bool on) { opt_elision_ = on; }
    void set_use_graph(bool on) { opt_graph_ = on; }
    void set_l2_hints(bool on) { opt_l2_hints_ = on; level_graph_key_ = 0; }
    // Levelised runs of this simulator need not keep rows after their last reader, except the rows of `keep_wires` (the batch
    // runner's designated outputs): dead rows are dropped from L2 (discard.global.L2) before they are written back to HBM.
    // After such a run only the kept wires (and rows read two or more levels above their gate) may be read back.
    void set_discard(bool on, const uint32_t *keep_wires, uint32_t n);
    void set_budget_inclusive(bool on) { opt_budget_inclusive_ = on; }
    void set_use_frontier(bool on) { opt_frontier_ = on; frontier_valid_ = false; }
    void set_device_loop(bool on) { opt_device_loop_ = on; }
    void set_period2_skip(bool on) { opt_period2_ = on; }
    void set_pdl(uint32_t mode) { opt_pdl_ = mode <= 2 ? mode : 0; level_graph_key_ = 0; }
    // Level-pair fusion (pair.cuh): acts only on levelised runs with set_discard on, for netlists of one/two-input gates.
    // 0 off, 1 on, 2 (default) automatic: on up to kPairAutoWidth gates per level on average.  Measured (one B200, final kernels):
    // 32 gates per level (config 2) +38 %, 5000 (config 3) +5 %, 50 000 (config 5) -2 %.
    void set_pair_fusion(uint32_t mode) { opt_pair_ = mode <= 2 ? mode : 2; level_graph_key_ = 0; }
    static constexpr uint32_t kPairAutoWidth = 16384;
    void set_direct_cycle(bool on) { opt_direct_ = on; }
    // Unit-delay work-list kernels keep the (at most four) rows with the largest fan-out (>= 100 gates) in shared memory.
    void set_hot_staging(bool on) { opt_hot_ = on; }
    void hot_rows(uint32_t *rows, uint32_t *n);  // store rows staged (n = 0: none qualifies, or the option is off)
    // launches of the last levelised pass / gates evaluated as children / parent rows never stored (0 when the plan is off)
    void pair_info(uint64_t *launches, uint64_t *children, uint64_t *unstored) const;
    // SURVEY.md A.6-2 alternative (agreeing valid drivers do not conflict); refused for netlists with two firm drivers on a net
    bool set_lenient_conflicts(bool on) {
        if (on && net_.multi_driver) return false;
        opt_lenient_ = on;
        return true;
    }
    void set_eval_unroll(uint32_t u) { opt_u_ = (u == 2 || u == 4) ? u : 0; level_graph_key_ = 0; }

   private:
    int ensure_device();
    int ensure_lanes();
    int ensure_second_buffer();
    int ensure_stage(bool out, size_t words_per_plane);
    int drives_begin();   // main stream is about to touch the drive store
    int drives_end();     // ... and is done with it
    int ensure_rows_buffer(uint32_t n);
    void free_store();
    int drive_target(uint32_t bitnet, uint64_t **val, uint64_t **vld);  // row pointers for a bit-net's drive
    int upload_rows(const uint32_t *wires, uint32_t n, bool drive_rows);
    int run_levelised(bool async);
    int run_jacobi(uint64_t max_steps);
    int run_device_loop(uint64_t &k, uint64_t last, bool &finished, int &result);
    int run_resident(uint64_t max_steps, bool levelised, const void *inline_bits = nullptr, uint32_t n_inline = 0);
    int pending_to_bits(void *out, uint32_t cap, uint32_t *n);
    bool direct_ready();
    int materialise_valid();
    uint32_t resident_wpc() const;

    std::shared_ptr<NetShare> share_;
    const Compiled &net_;
    std::shared_ptr<DeviceNet> dnet_;
    int device_ = -1;
    bool device_ready_ = false;
    void *stream_ = nullptr;  // cudaStream_t
    bool own_stream_ = false, stream_set_ = false;

    uint64_t n_lanes_ = 0;
    uint32_t T_ = 0, stride_ = 0;
    uint64_t store_bytes_ = 0;

    // netlist on the device (aliases into dnet_, which owns them)
    uint32_t *d_fan0_ = nullptr, *d_fan1_ = nullptr, *d_fan2_ = nullptr, *d_wide_off_ = nullptr, *d_wide_in_ = nullptr;
    uint8_t *d_wide_aux_ = nullptr;
    uint8_t *d_kind2_ = nullptr, *d_far2_ = nullptr, *d_wide_kind_ = nullptr, *d_wide_nsel_ = nullptr;

    // wire-state store: two buffers (second one only for unit-delay runs), separate planes
    uint64_t *val_[2] = {nullptr, nullptr}, *vld_[2] = {nullptr, nullptr};
    int cur_ = 0;
    // drive store for source rows [n_src][stride]
    uint64_t *drv_val_ = nullptr, *drv_vld_ = nullptr;
    // drives on gate-driven bits (conflict-prone, rare): slot per bit-net
    std::unordered_map<uint32_t, uint32_t> extra_slot_;
    std::vector<uint32_t> extra_rows_;
    uint32_t extra_cap_ = 0;
    uint64_t *xdrv_val_ = nullptr, *xdrv_vld_ = nullptr;
    uint32_t *d_extra_rows_ = nullptr;
    bool extra_rows_dirty_ = false;
    uint32_t n_plain_extra_ = 0;  // slots on plain gate rows (merged by k_extra_drive); the others belong to tri-state nets
    // tri-state nets: resolve tables on the device; entry -> slot of its external drive (or -1)
    uint32_t *d_res_off_ = nullptr, *d_res_in_ = nullptr;
    int32_t *d_res_xslot_ = nullptr;
    Level *d_levels_ = nullptr;
    std::vector<int32_t> res_xslot_;
    bool res_xslot_dirty_ = false;
    int sync_extra_tables();  // uploads d_extra_rows_ / d_res_xslot_ when they changed

    // per-lane status words [T]
    uint64_t *d_changed_ = nullptr, *d_conf_step_ = nullptr, *d_live_ = nullptr, *d_conflict_ = nullptr, *d_unsettled_ = nullptr;
    RunFlags *d_flags_ = nullptr, *h_flags_ = nullptr;
    void *h_flags_dev_ = nullptr;     // h_flags_ as the device sees it (mapped page-locked memory), or nullptr
    uint8_t *h_snap_ = nullptr, *h_snap_dev_ = nullptr;  // page-locked lane-0 snapshot the one-CTA resident run stores into
    size_t h_snap_cap_ = 0;
    bool opt_direct_ = true;          // one-CTA resident runs: drives in the kernel arguments, flags and snapshot stored to the host

    // staging for host <-> device plane transfers, on their own copy streams so that the H2D of
    // the next tile and the D2H of the previous one overlap the settle of the current tile
    uint64_t *stage_val_[2] = {nullptr, nullptr}, *stage_vld_[2] = {nullptr, nullptr};  // [0] in, [1] out
    size_t stage_words_[2] = {0, 0};
    void *cin_stream_ = nullptr, *cout_stream_ = nullptr;                  // cudaStream_t
    void *ev_drive_consumed_ = nullptr, *ev_drives_ready_ = nullptr;       // cudaEvent_t
    void *ev_gathered_ = nullptr, *ev_out_free_ = nullptr;
    void *side_stream_ = nullptr, *ev_settled_ = nullptr, *ev_side_done_ = nullptr;  // gather_wires_multi
    bool side_pending_ = false;
    bool cin_pending_ = false;
    void *ev_flags_[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t *d_rows_ = nullptr;
    uint32_t rows_cap_ = 0;
    uint64_t rows_key_ = 0, ptrs_key_ = 0;  // hashes of the cached wire lists (0 = none)
    std::vector<uint32_t> rows_wires_, ptrs_wires_;  // the cached lists themselves (a hash hit is confirmed against them)
    uint8_t *d_pack_ = nullptr;
    size_t pack_cap_ = 0;
    // set_wire_drive queue (flushed before the next run) and the lane-0 snapshot get_wire_state reads
    struct PendingDrive {
        uint32_t wire, p0[8], p1[8];
    };
    std::vector<PendingDrive> pending_;
    void *d_pend_ = nullptr;
    size_t pend_cap_ = 0;
    std::vector<uint8_t> snap_;  // [2][n_rows / 8 + 1]: value bits, then valid bits
    bool snap_valid_ = false;
    int flush_pending_drives();
    int refresh_snapshot();

    bool cold_ = true;
    // valid-plane elision: d_kv_[row] = 1 while the row's valid plane is all-ones and not stored
    uint8_t *d_kv_ = nullptr;
    uint32_t *d_all_valid_ = nullptr;  // 1 when every source row of the last elided run was valid
    bool elided_ = false, opt_elision_ = false;
    // CUDA graph of the level sweep (levelised path)
    void *level_graph_exec_ = nullptr;  // cudaGraphExec_t
    uint64_t level_graph_key_ = 0;
    uint32_t level_graph_nodes_ = 0;
    bool opt_graph_ = true;
    // unit-delay changed-row frontier
    uint8_t *d_rowchg_[2] = {nullptr, nullptr};
    uint32_t *d_act_list_[4] = {nullptr, nullptr, nullptr, nullptr};  // compacted work lists: one/two-input gates, three-input gates, rows to refresh, wide gates
    size_t rowchg_bytes() const { return (((size_t)net_.n_rows + 15) & ~(size_t)15) + 16; }  // flags, padding, the list counters
    // device-side run_sim loop (k_unit_delay_loop): CTAs of one cooperative launch, result block
    uint32_t loop_ctas_ = 0;
    struct LoopResult *d_loop_result_ = nullptr;
    void *h_loop_result_ = nullptr;
    bool opt_device_loop_ = true;
    int sm_count_ = 148;
    int occ_list_[4] = {3, 2, 4, 4};  // resident CTAs per SM of k_eval_list<2>, k_eval_list<3>, k_copy_list, k_eval_wide_list
    int chg_w_ = 0;
    bool frontier_valid_ = false, opt_frontier_ = true;
    bool opt_lenient_ = false;
    bool opt_budget_inclusive_ = false;  // SURVEY.md A.6-1 switch: `steps >= max_steps` instead of `>`
    uint32_t opt_u_ = 0;  // 0 = automatic gates-per-thread choice in k_eval2
    bool opt_discard_ = false;
    struct RowRange { uint32_t begin, end; };
    std::vector<RowRange> disc_short_, disc_never_;  // per level: rows of the one/two-input class that may be dropped
    bool opt_l2_hints_ = false;
    uint32_t opt_pair_ = 2;  // level-pair fusion: 0 off, 1 on, 2 automatic (by the average level width)
    bool opt_hot_ = false;  // k_eval_list stages the hottest fan-in rows in shared memory (measured: L2 serves them as well)
    bool hot_ready_ = false;
    uint32_t n_hot_ = 0, hot_rows_[4] = {0, 0, 0, 0};
    void find_hot_rows();
    std::vector<uint32_t> keep_rows_;      // sorted rows kept by set_discard
    std::shared_ptr<PairPlanDev> pair_;    // null: not built yet, or the netlist is not eligible (pair_tried_)
    bool pair_tried_ = false, last_pair_ = false;
    int ensure_pair_plan();
    uint32_t opt_pdl_ = 2;  // programmatic dependent launch between the level kernels: 0 off, 1 early release, 2 late release  // levelised pass: evict-first hint on fan-ins produced long ago (Compiled::far2)
    bool opt_period2_ = true;   // unit-delay loops fast-forward over an exact period-2 oscillation
    uint32_t opt_flags_ = 0;    // SemFlags: the reconstructed points of SURVEY.md A.6-4..7, one bit each
    bool use_resident_ = true, resident_attr_set_ = false;
    uint32_t last_mode_ = 0;
    uint64_t last_apps_ = 0;
};

}  // namespace b2
