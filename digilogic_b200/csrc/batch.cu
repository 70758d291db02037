// batch.cu — the batch runner behind b2_batch_* and the communicator behind b2_comm_* (see batch.hpp, sweep.cuh).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and enums only: every NCCL entry point is resolved with dlsym at run time

#include <algorithm>
#include <cstring>
#include <string>

#include "../../include/digilogic_b200.h"
#include "batch.hpp"
#include "sweep.cuh"

namespace b2 {

void count_launch();  // engine.cu: b2_kernel_launches()

#define CU_TRY(expr)                                                                      \
    do {                                                                                  \
        cudaError_t e_ = (expr);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            set_last_error(std::string(#expr) + ": " + cudaGetErrorString(e_));           \
            return e_ == cudaErrorMemoryAllocation ? B2_ERR_OUT_OF_MEMORY : B2_ERR_CUDA; \
        }                                                                                 \
    } while (0)

#define ST ((cudaStream_t)stream_)

// ---------------------------------------------------------------------------- NCCL, resolved at run time

namespace {
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string why;
};

NcclApi &nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    // the copy a host process has already loaded (e.g. the one PyTorch ships) wins, so that one process never runs two NCCLs
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        api.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
        return api;
    }
    auto sym = [&](const char *name) {
        void *p = dlsym(h, name);
        if (!p) api.why = std::string("NCCL symbol missing: ") + name;
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce && api.GetErrorString;
    return api;
}

int nccl_check(ncclResult_t r, const char *what) {
    if (r == ncclSuccess) return B2_OK;
    set_last_error(std::string(what) + ": " + nccl().GetErrorString(r));
    return B2_ERR_CUDA;
}
}  // namespace

int Comm::unique_id(uint8_t id[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    NcclApi &n = nccl();
    if (!n.ok) {
        set_last_error(n.why);
        return B2_ERR_NO_DEVICE;
    }
    ncclUniqueId u;
    int rc = nccl_check(n.GetUniqueId(&u), "ncclGetUniqueId");
    if (rc) return rc;
    std::memcpy(id, &u, 128);
    return B2_OK;
}

int Comm::create(Comm **out, int n_ranks, int rank, const uint8_t id[128], int device) {
    NcclApi &n = nccl();
    if (!n.ok) {
        set_last_error(n.why);
        return B2_ERR_NO_DEVICE;
    }
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return B2_ERR_INVALID_ARG;
    CU_TRY(cudaSetDevice(device));
    ncclUniqueId u;
    std::memcpy(&u, id, 128);
    ncclComm_t c = nullptr;
    int rc = nccl_check(n.CommInitRank(&c, n_ranks, u, rank), "ncclCommInitRank");
    if (rc) return rc;
    Comm *cm = new Comm();
    cm->comm_ = c;
    cm->n_ranks_ = n_ranks;
    cm->rank_ = rank;
    cm->device_ = device;
    *out = cm;
    return B2_OK;
}

Comm::~Comm() {
    if (comm_) nccl().CommDestroy((ncclComm_t)comm_);
}

int Comm::all_gather_u64(uint64_t *buf, size_t count, void *stream) {
    return nccl_check(nccl().AllGather(buf + (size_t)rank_ * count, buf, count, ncclUint64, (ncclComm_t)comm_, (cudaStream_t)stream),
                      "ncclAllGather");
}

int Comm::all_reduce_max_u32(uint32_t *buf, size_t count, void *stream) {
    return nccl_check(nccl().AllReduce(buf, buf, count, ncclUint32, ncclMax, (ncclComm_t)comm_, (cudaStream_t)stream), "ncclAllReduce");
}

// ---------------------------------------------------------------------------- Batch: setup

__global__ void k_status_max(uint32_t *status, uint32_t v) { atomicMax(status, v); }

Batch::Batch(const Engine &proto, const uint32_t *inputs, uint32_t n_in, const uint32_t *outputs, uint32_t n_out, uint32_t tile_words,
             uint32_t n_slots)
    : share_(proto.share()), net_(share_->net), inputs_(inputs, inputs + n_in), outputs_(outputs, outputs + n_out) {
    proto_.reset(new Engine(share_));
    proto_->copy_options_from(proto);
    device_ = proto.device();
    T_ = std::max<uint32_t>(tile_words, 2);
    T_ += T_ & 1;  // whole 128-bit pairs of lane-words
    K_ = std::max<uint32_t>(n_slots, 1);
    // The level sweep computes the unique fixed point of an acyclic single-driver netlist whose inputs are undriven nets
    // (a drive on a gate output would have to be merged and checked for conflicts per lane: the replicas do that).
    sweep_eligible_ = !net_.cyclic && !net_.multi_driver && T_ <= 512;
    for (uint32_t w : inputs_)
        if (w < net_.n_wires() && net_.bit_driven[net_.bit_off[w]]) sweep_eligible_ = false;
}

int Batch::validate() const {
    for (const std::vector<uint32_t> *list : {&inputs_, &outputs_})
        for (uint32_t w : *list) {
            if (w >= net_.n_wires()) return B2_ERR_INVALID_WIRE_ID;
            if (net_.width[w] != 1) return B2_ERR_WIRE_WIDTH_INCOMPATIBLE;  // lane planes carry one bit per wire
        }
    return B2_OK;
}

Batch::~Batch() {
    if (device_ < 0 || cudaSetDevice(device_) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    if (stream_set_) cudaStreamSynchronize(ST);
    release_registered();
    replicas_.clear();
    for (void *s : rstreams_) cudaStreamDestroy((cudaStream_t)s);
    for (void *e : revents_) cudaEventDestroy((cudaEvent_t)e);
    if (ev_fork_) cudaEventDestroy((cudaEvent_t)ev_fork_);
    for (size_t r = 0; r < peer_val_.size(); r++) {
        if (peer_val_[r] && peer_val_[r] != gat_val_) cudaIpcCloseMemHandle(peer_val_[r]);
        if (peer_vld_[r] && peer_vld_[r] != gat_vld_) cudaIpcCloseMemHandle(peer_vld_[r]);
    }
    cudaFree(gat_val_);
    cudaFree(gat_vld_);
    cudaFree(val_);
    cudaFree(vld_);
    cudaFree(d_segs_);
    cudaFree(d_seg_task0_);
    cudaFree(d_in_rows_);
    cudaFree(d_out_rows_);
    cudaFree(d_sync_);
    cudaFree(d_status_);
    if (h_status_) cudaFreeHost(h_status_);
    if (own_stream_ && stream_) cudaStreamDestroy(ST);
    dnet_.reset();
    cudaGetLastError();
}

int Batch::set_device(int device) {
    if (device_ready_) return device == device_ ? B2_OK : B2_ERR_INVALID_ARG;
    device_ = device;
    return B2_OK;
}

int Batch::set_stream(void *stream) {
    if (device_ready_ && stream_set_ && stream != stream_) CU_TRY(cudaStreamSynchronize(ST));
    if (own_stream_ && stream_) {
        cudaStreamDestroy(ST);
        own_stream_ = false;
    }
    stream_ = stream;
    stream_set_ = true;
    return B2_OK;
}

int Batch::set_option(int option, int64_t value) {
    switch (option) {
        case BATCH_OPT_SWEEP: opt_sweep_ = value != 0; return B2_OK;
        case BATCH_OPT_TASK_ITERS:
            if (value < 1 || value > 64) return B2_ERR_INVALID_ARG;
            opt_iters_ = (uint32_t)value;
            return B2_OK;
        case BATCH_OPT_UNROLL:
            if (value != 2 && value != 4) return B2_ERR_INVALID_ARG;
            opt_unroll_ = (uint32_t)value;
            return B2_OK;
        case BATCH_OPT_CTAS_PER_SM:
            if (value < 0 || value > 8) return B2_ERR_INVALID_ARG;
            opt_ctas_per_sm_ = (uint32_t)value;
            return B2_OK;
        case BATCH_OPT_AUTO_GATHER: opt_auto_gather_ = value != 0; return B2_OK;
        case BATCH_OPT_DISCARD:
            opt_discard_ = value != 0;
            for (auto &e : replicas_) e->set_discard(opt_discard_, outputs_.data(), (uint32_t)outputs_.size());
            return B2_OK;
        default: return B2_ERR_INVALID_ARG;
    }
}

int Batch::ensure_device() {
    if (device_ready_) {
        CU_TRY(cudaSetDevice(device_));
        return B2_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_last_error(std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                       " (this engine has no CPU path)");
        cudaGetLastError();
        return B2_ERR_NO_DEVICE;
    }
    if (device_ < 0 && cudaGetDevice(&device_) != cudaSuccess) device_ = 0;
    if (device_ >= count) {
        set_last_error("device ordinal out of range");
        return B2_ERR_INVALID_ARG;
    }
    CU_TRY(cudaSetDevice(device_));
    CU_TRY(cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, device_));
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, k_sweep<4>);
    if (e != cudaSuccess) {
        set_last_error(std::string("sm_100a kernels not loadable on this device: ") + cudaGetErrorString(e));
        cudaGetLastError();
        return B2_ERR_NO_DEVICE;
    }
    if (!stream_set_) {
        cudaStream_t s;
        CU_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        stream_ = s;
        own_stream_ = stream_set_ = true;
    }
    int rc = B2_OK;
    dnet_ = share_->acquire(device_, rc);
    if (!dnet_) return rc;
    if (!d_status_) {
        CU_TRY(cudaMalloc((void **)&d_status_, 16));
        CU_TRY(cudaMemset(d_status_, 0, 16));
        CU_TRY(cudaMallocHost((void **)&h_status_, 16));
        h_status_[0] = 0;
    }
    device_ready_ = true;
    return B2_OK;
}

template <typename Tv>
static int upload_vec(Tv **dst, const std::vector<Tv> &src) {
    cudaFree(*dst);
    *dst = nullptr;
    CU_TRY(cudaMalloc((void **)dst, std::max<size_t>(src.size(), 1) * sizeof(Tv)));
    if (!src.empty()) CU_TRY(cudaMemcpy(*dst, src.data(), src.size() * sizeof(Tv), cudaMemcpyHostToDevice));
    return B2_OK;
}

// Slots, the task table of one pass, the ticket / slot counters.
int Batch::ensure_sweep() {
    if (sweep_ready_ && built_iters_ == opt_iters_ && built_unroll_ == opt_unroll_) return B2_OK;
    int rc = ensure_device();
    if (rc) return rc;
    CU_TRY(cudaStreamSynchronize(ST));
    const uint32_t pairs = T_ / 2;
    uint32_t bx = 1;
    while (bx < pairs) bx <<= 1;
    const uint32_t by = 256 / bx;
    const uint32_t per2 = by * opt_unroll_ * opt_iters_, per3 = by * 2 * opt_iters_, perw = by * opt_iters_;
    gates_per_task_ = per2;
    std::vector<SweepSeg> segs;
    uint32_t task0 = 0;
    auto add = [&](uint32_t type, uint32_t begin, uint32_t end, uint32_t per, uint32_t need) {
        if (end <= begin) return;
        SweepSeg s{type, begin, end, per, (end - begin + per - 1) / per, task0, need, 0};
        task0 += s.n_tasks;
        segs.push_back(s);
    };
    // stimulus and output tasks move 16 B per row-word: about the bytes of a gate task
    const uint32_t per_io = std::max<uint32_t>(by * 8, 8);
    add(SEG_STIM, 0, (uint32_t)inputs_.size(), per_io, 0);
    for (const Level &lv : net_.levels) {
        const uint32_t need = task0;  // everything in front of this level
        add(SEG_G2, lv.g2_begin, lv.g2_end, per2, need);
        add(SEG_G3, lv.g3_begin, lv.g3_end, per3, need);
        add(SEG_WIDE, lv.wide_begin, lv.wide_end, perw, need);
    }
    add(SEG_OUT, 0, (uint32_t)outputs_.size(), per_io, task0);
    if (segs.empty()) {  // nothing to do at all: one empty stimulus segment keeps the decode well-defined
        segs.push_back(SweepSeg{SEG_STIM, 0, 0, 1, 1, 0, 0, 0});
        task0 = 1;
    }
    n_segs_ = (uint32_t)segs.size();
    tasks_per_pass_ = task0;
    std::vector<uint32_t> t0(n_segs_ + 1);
    for (uint32_t i = 0; i < n_segs_; i++) t0[i] = segs[i].task0;
    t0[n_segs_] = tasks_per_pass_;
    if ((rc = upload_vec(&d_segs_, segs)) || (rc = upload_vec(&d_seg_task0_, t0))) return rc;
    if (!sweep_ready_) {
        std::vector<uint32_t> rows(inputs_.size());
        for (size_t i = 0; i < inputs_.size(); i++) rows[i] = net_.row_of_bit[net_.bit_off[inputs_[i]]];
        if ((rc = upload_vec(&d_in_rows_, rows))) return rc;
        rows.resize(outputs_.size());
        for (size_t i = 0; i < outputs_.size(); i++) rows[i] = net_.row_of_bit[net_.bit_off[outputs_[i]]];
        if ((rc = upload_vec(&d_out_rows_, rows))) return rc;
        // the stores: sources that are not inputs stay High-Z (0, 0) for good; every other row is rewritten by each pass
        const size_t words = (size_t)std::max<uint32_t>(net_.n_rows, 1) * T_ * K_;
        CU_TRY(cudaMalloc((void **)&val_, words * 8));
        CU_TRY(cudaMalloc((void **)&vld_, words * 8));
        CU_TRY(cudaMemsetAsync(val_, 0, words * 8, ST));
        CU_TRY(cudaMemsetAsync(vld_, 0, words * 8, ST));
        CU_TRY(cudaMalloc((void **)&d_sync_, (size_t)(K_ + 1) * 128));
        store_bytes_ += words * 16;
    }
    int occ = 0;
    if (opt_unroll_ == 4) CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep<4>, 256, 0));
    else CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep<2>, 256, 0));
    occ = std::max(occ, 1);
    if (opt_ctas_per_sm_) occ = std::min<int>(occ, (int)opt_ctas_per_sm_);
    grid_ = (uint32_t)(occ * sm_count_);
    built_iters_ = opt_iters_;
    built_unroll_ = opt_unroll_;
    sweep_ready_ = true;
    return B2_OK;
}

int Batch::ensure_replicas() {
    if (!replicas_.empty()) return B2_OK;
    int rc = ensure_device();
    if (rc) return rc;
    cudaEvent_t ev;
    CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ev_fork_ = ev;
    for (uint32_t k = 0; k < K_; k++) {
        cudaStream_t s;
        CU_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        rstreams_.push_back(s);
        CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        revents_.push_back(ev);
        std::unique_ptr<Engine> e(new Engine(share_));
        e->copy_options_from(*proto_);
        e->set_discard(opt_discard_, outputs_.data(), (uint32_t)outputs_.size());
        if ((rc = e->set_device(device_)) || (rc = e->set_stream(s)) || (rc = e->set_lanes((uint64_t)T_ * 64))) return rc;
        store_bytes_ += e->store_bytes();
        replicas_.push_back(std::move(e));
    }
    return B2_OK;
}

int Batch::info(BatchInfo *out) {
    std::memset(out, 0, sizeof(*out));
    out->tile_words = T_;
    out->n_slots = K_;
    out->sweep = sweep_eligible_ && opt_sweep_ ? 1 : 0;
    out->tasks_per_pass = tasks_per_pass_;
    out->grid_ctas = grid_;
    out->gates_per_task = gates_per_task_;
    out->n_ranks = comm_ ? (uint32_t)comm_->n_ranks() : 1;
    out->rank = comm_ ? (uint32_t)comm_->rank() : 0;
    out->peers_mapped = peers_mapped_ ? 1 : 0;
    if (!replicas_.empty()) {
        uint64_t pl = 0;
        replicas_[0]->pair_info(&pl, nullptr, nullptr);
        out->pair_launches = (uint32_t)pl;
    }
    out->store_bytes = store_bytes_;
    out->words_local = words_local_;
    out->last_tiles = last_tiles_;
    out->last_tickets = last_tickets_;
    return B2_OK;
}

// ---------------------------------------------------------------------------- Batch: multi-GPU

int Batch::bind_comm(Comm *comm, uint64_t words_local) {
    if (!comm || words_local == 0) return B2_ERR_INVALID_ARG;
    if (comm_) return B2_ERR_INVALID_ARG;
    if (device_ >= 0 && device_ != comm->device() && device_ready_) return B2_ERR_INVALID_ARG;
    device_ = comm->device();
    int rc = ensure_device();
    if (rc) return rc;
    const size_t words = (size_t)comm->n_ranks() * std::max<size_t>(outputs_.size(), 1) * words_local;
    CU_TRY(cudaMalloc((void **)&gat_val_, words * 8));
    CU_TRY(cudaMalloc((void **)&gat_vld_, words * 8));
    CU_TRY(cudaMemsetAsync(gat_val_, 0, words * 8, ST));
    CU_TRY(cudaMemsetAsync(gat_vld_, 0, words * 8, ST));
    CU_TRY(cudaStreamSynchronize(ST));
    store_bytes_ += words * 16;
    comm_ = comm;
    words_local_ = words_local;
    peer_val_.assign(comm->n_ranks(), nullptr);
    peer_vld_.assign(comm->n_ranks(), nullptr);
    peer_val_[comm->rank()] = gat_val_;
    peer_vld_[comm->rank()] = gat_vld_;
    return B2_OK;
}

int Batch::export_gather_handle(uint8_t handle[128]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!comm_) return B2_ERR_INVALID_ARG;
    CU_TRY(cudaSetDevice(device_));
    cudaIpcMemHandle_t hv, hm;
    CU_TRY(cudaIpcGetMemHandle(&hv, gat_val_));
    CU_TRY(cudaIpcGetMemHandle(&hm, gat_vld_));
    std::memcpy(handle, &hv, 64);
    std::memcpy(handle + 64, &hm, 64);
    return B2_OK;
}

int Batch::import_gather_handles(const uint8_t *handles) {
    if (!comm_ || !handles) return B2_ERR_INVALID_ARG;
    if (peers_mapped_) return B2_OK;
    CU_TRY(cudaSetDevice(device_));
    for (int r = 0; r < comm_->n_ranks(); r++) {
        if (r == comm_->rank()) continue;
        cudaIpcMemHandle_t hv, hm;
        std::memcpy(&hv, handles + (size_t)r * 128, 64);
        std::memcpy(&hm, handles + (size_t)r * 128 + 64, 64);
        void *pv = nullptr, *pm = nullptr;
        CU_TRY(cudaIpcOpenMemHandle(&pv, hv, cudaIpcMemLazyEnablePeerAccess));
        peer_val_[r] = (uint64_t *)pv;
        CU_TRY(cudaIpcOpenMemHandle(&pm, hm, cudaIpcMemLazyEnablePeerAccess));
        peer_vld_[r] = (uint64_t *)pm;
    }
    peers_mapped_ = true;
    return B2_OK;
}

int Batch::gathered(uint64_t **val, uint64_t **vld) {
    if (!comm_) return B2_ERR_INVALID_ARG;
    *val = gat_val_;
    *vld = gat_vld_;
    return B2_OK;
}

// Completes the gather of the last run's output planes: with the peers' planes mapped the rows are already there (the output
// tasks stored them over NVLink) and only the ranks have to meet; otherwise NCCL moves them.  The MAX all-reduce of the run
// result is that meeting point.
int Batch::gather_outputs() {
    if (!comm_) return B2_ERR_INVALID_ARG;
    int rc = ensure_device();
    if (rc) return rc;
    if (!peers_mapped_ && comm_->n_ranks() > 1) {
        const size_t count = outputs_.size() * words_local_;
        if ((rc = comm_->all_gather_u64(gat_val_, count, ST)) || (rc = comm_->all_gather_u64(gat_vld_, count, ST))) return rc;
    }
    if (comm_->n_ranks() > 1 && (rc = comm_->all_reduce_max_u32(d_status_, 1, ST))) return rc;
    gather_pending_ = false;
    return B2_OK;
}

// ---------------------------------------------------------------------------- Batch: runs

// Device-usable address of a caller buffer: device memory as it is, page-locked host memory through its mapping, pageable
// host memory page-locked for the duration of the run (released by wait()).
int Batch::device_pointer(const void *p, size_t bytes, void **dev, bool *is_host) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        at.type = cudaMemoryTypeUnregistered;
    }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) {
        *dev = const_cast<void *>(p);
        *is_host = false;
        return B2_OK;
    }
    *is_host = true;
    if (at.type == cudaMemoryTypeUnregistered) {
        CU_TRY(cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
        registered_.push_back(const_cast<void *>(p));
    }
    CU_TRY(cudaHostGetDevicePointer(dev, const_cast<void *>(p), 0));
    return B2_OK;
}

void Batch::release_registered() {
    for (void *p : registered_) cudaHostUnregister(p);
    registered_.clear();
}

int Batch::run(int stim_mode, uint64_t seed, uint64_t word_begin, const uint64_t *in_val, const uint64_t *in_vld, uint64_t in_stride,
               uint64_t in_ring, uint64_t n_words, uint64_t max_steps, uint64_t *out_val, uint64_t *out_vld, uint64_t out_stride,
               uint64_t out_ring) {
    if (stim_mode < 0 || stim_mode > 2) return B2_ERR_INVALID_ARG;
    if (stim_mode == 2 && (!in_val || !in_vld)) return B2_ERR_NULL;
    if ((out_val == nullptr) != (out_vld == nullptr)) return B2_ERR_NULL;
    if (n_words == 0) return B2_OK;
    if (stim_mode == 2 && in_stride < (in_ring ? in_ring : n_words)) return B2_ERR_RANGE;
    if (out_val && out_stride < (out_ring ? out_ring : n_words)) return B2_ERR_RANGE;
    if ((in_ring && in_ring % T_) || (out_ring && out_ring % T_)) return B2_ERR_INVALID_ARG;  // tiles never straddle a ring's end
    if (comm_ && n_words > words_local_) return B2_ERR_RANGE;
    int rc = validate();
    if (rc) return rc;
    if ((rc = ensure_device())) return rc;
    // the ranks meet before a run overwrites gathered planes a peer may still be reading
    if (comm_ && comm_->n_ranks() > 1 && peers_mapped_ && (rc = comm_->all_reduce_max_u32(d_status_ + 2, 1, ST))) return rc;
    const bool dag_ok = max_steps >= (uint64_t)net_.depth + 1;
    if (sweep_eligible_ && opt_sweep_ && dag_ok)
        rc = run_sweep(stim_mode, seed, word_begin, in_val, in_vld, in_stride, in_ring, n_words, out_val, out_vld, out_stride, out_ring);
    else
        rc = run_replicas(stim_mode, seed, word_begin, in_val, in_vld, in_stride, in_ring, n_words, max_steps, out_val, out_vld,
                          out_stride, out_ring);
    if (rc) return rc;
    gather_pending_ = comm_ != nullptr;
    if (gather_pending_ && opt_auto_gather_) return gather_outputs();
    return B2_OK;
}

int Batch::run_sweep(int stim_mode, uint64_t seed, uint64_t word_begin, const uint64_t *in_val, const uint64_t *in_vld,
                     uint64_t in_stride, uint64_t in_ring, uint64_t n_words, uint64_t *out_val, uint64_t *out_vld, uint64_t out_stride,
                     uint64_t out_ring) {
    int rc = ensure_sweep();
    if (rc) return rc;
    SweepArgs a;
    std::memset(&a, 0, sizeof(a));
    a.fan0 = dnet_->fan0; a.fan1 = dnet_->fan1; a.fan2 = dnet_->fan2; a.kind2 = dnet_->kind2;
    a.wide_off = dnet_->wide_off; a.wide_in = dnet_->wide_in; a.wide_kind = dnet_->wide_kind; a.wide_nsel = dnet_->wide_nsel; a.wide_aux = dnet_->wide_aux;
    a.val = val_; a.vld = vld_;
    a.slot_words = (uint64_t)std::max<uint32_t>(net_.n_rows, 1) * T_;
    a.segs = d_segs_; a.seg_task0 = d_seg_task0_;
    a.n_segs = n_segs_; a.tasks_per_pass = tasks_per_pass_;
    a.K = K_; a.T = T_; a.pairs = T_ / 2;
    a.n_src = net_.n_src; a.n_fast = net_.n_fast();
    a.sem = net_.sem_flags;
    a.n_words = n_words;
    a.n_tiles = (n_words + T_ - 1) / T_;
    a.ticket = d_sync_;
    a.done = d_sync_ + 16;
    const uint64_t groups = (a.n_tiles + K_ - 1) / K_;
    a.total_tickets = groups * K_ * tasks_per_pass_;
    a.in_rows = d_in_rows_;
    a.stim_mode = stim_mode;
    a.seed = seed;
    a.word_begin = word_begin;
    bool is_host = false;
    if (stim_mode == 2) {
        void *dv = nullptr, *dm = nullptr;
        const size_t bytes = ((inputs_.size() ? inputs_.size() - 1 : 0) * in_stride + (in_ring ? in_ring : n_words)) * 8;
        if ((rc = device_pointer(in_val, bytes, &dv, &is_host)) || (rc = device_pointer(in_vld, bytes, &dm, &is_host))) return rc;
        a.in_val = (const uint64_t *)dv;
        a.in_vld = (const uint64_t *)dm;
        a.in_stride = in_stride;
        a.in_ring = in_ring;
        a.in_vec = (((uintptr_t)dv | (uintptr_t)dm) % 16 == 0 && in_stride % 2 == 0) ? 1 : 0;
    }
    a.out_rows = d_out_rows_;
    a.n_dst = 0;
    if (out_val) {
        void *dv = nullptr, *dm = nullptr;
        const size_t bytes = ((outputs_.size() ? outputs_.size() - 1 : 0) * out_stride + (out_ring ? out_ring : n_words)) * 8;
        if ((rc = device_pointer(out_val, bytes, &dv, &is_host)) || (rc = device_pointer(out_vld, bytes, &dm, &is_host))) return rc;
        SweepDst &d = a.dst[a.n_dst++];
        d.val = (uint64_t *)dv; d.vld = (uint64_t *)dm;
        d.stride = out_stride; d.word0 = 0; d.ring = out_ring;
        d.vec = (((uintptr_t)dv | (uintptr_t)dm) % 16 == 0 && out_stride % 2 == 0) ? 1 : 0;
    }
    if (comm_) {
        const uint64_t plane = (uint64_t)outputs_.size() * words_local_;  // one rank's share of a gathered plane
        for (int r = 0; r < comm_->n_ranks(); r++) {
            if (!peer_val_[r]) continue;  // peers not mapped: gather_outputs() moves the rows with NCCL
            if (a.n_dst == kMaxDst) return B2_ERR_RANGE;
            SweepDst &d = a.dst[a.n_dst++];
            d.val = peer_val_[r] + (uint64_t)comm_->rank() * plane;
            d.vld = peer_vld_[r] + (uint64_t)comm_->rank() * plane;
            d.stride = words_local_; d.word0 = 0; d.ring = 0;
            d.vec = words_local_ % 2 == 0 ? 1 : 0;
        }
    }
    if (outputs_.empty()) a.n_dst = 0;
    CU_TRY(cudaMemsetAsync(d_sync_, 0, (size_t)(K_ + 1) * 128, ST));
    const uint32_t pairs = T_ / 2;
    uint32_t bx = 1;
    while (bx < pairs) bx <<= 1;
    const dim3 block(bx, 256 / bx);
    const uint32_t grid = (uint32_t)std::min<uint64_t>(grid_, a.total_tickets);
    if (opt_unroll_ == 4) k_sweep<4><<<grid, block, 0, ST>>>(a);
    else k_sweep<2><<<grid, block, 0, ST>>>(a);
    count_launch();
    CU_TRY(cudaGetLastError());
    last_tiles_ = a.n_tiles;
    last_tickets_ = a.total_tickets;
    return B2_OK;
}

// Netlists with feedback, tri-state nets or registers, drives on gate outputs, or a step budget below the depth: K simulator
// replicas (sharing the device copy of the netlist) on K streams take alternate tiles through Engine::run.
int Batch::run_replicas(int stim_mode, uint64_t seed, uint64_t word_begin, const uint64_t *in_val, const uint64_t *in_vld,
                        uint64_t in_stride, uint64_t in_ring, uint64_t n_words, uint64_t max_steps, uint64_t *out_val,
                        uint64_t *out_vld, uint64_t out_stride, uint64_t out_ring) {
    int rc = ensure_replicas();
    if (rc) return rc;
    if (n_words % T_) return B2_ERR_RANGE;  // whole tiles only on this path
    bool out_is_host = false;
    if (out_val) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, out_val) != cudaSuccess) {
            cudaGetLastError();
            at.type = cudaMemoryTypeUnregistered;
        }
        out_is_host = !(at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
    }
    CU_TRY(cudaEventRecord((cudaEvent_t)ev_fork_, ST));
    for (void *s : rstreams_) CU_TRY(cudaStreamWaitEvent((cudaStream_t)s, (cudaEvent_t)ev_fork_, 0));
    const uint64_t n_tiles = n_words / T_;
    const uint32_t n_in = (uint32_t)inputs_.size(), n_out = (uint32_t)outputs_.size();
    int worst = 0;
    for (uint64_t t = 0; t < n_tiles; t++) {
        Engine &e = *replicas_[t % K_];
        const uint64_t w0 = t * T_;
        // a slot is reused by later tiles, i.e. by other lanes: where state survives a run (feedback, registers) every tile
        // starts as a freshly built simulator — a batch run is one run_sim of L new simulators
        if (net_.cyclic) e.reset_cold();
        if (stim_mode < 2) {
            if ((rc = e.fill_stimulus(inputs_.data(), n_in, seed, word_begin + w0, stim_mode))) return rc;
        } else {
            const uint64_t w = in_ring ? w0 % in_ring : w0;
            if ((rc = e.set_drives_lanes(inputs_.data(), n_in, in_val + w, in_vld + w, in_stride))) return rc;
        }
        const int r = e.run(max_steps, nullptr, nullptr, true);
        if (r < 0) return -r;
        worst = std::max(worst, r);
        if (n_out) {
            // every device destination of the tile's output rows — the caller's planes, this rank's share of the gathered planes
            // on this GPU and on every peer whose planes are mapped — is served by ONE kernel on the replica's side stream
            // (rows read once, 128-bit stores, peers over NVLink) while the replica's stream goes on with its next tile
            std::vector<Engine::GatherDst> dsts;
            if (out_val) {
                const uint64_t w = out_ring ? w0 % out_ring : w0;
                if (out_is_host) {
                    if ((rc = e.gather_wires(outputs_.data(), n_out, out_val + w, out_vld + w, out_stride, false, true))) return rc;
                } else {
                    dsts.push_back(Engine::GatherDst{out_val + w, out_vld + w, out_stride});
                }
            }
            if (comm_) {
                const uint64_t plane = (uint64_t)n_out * words_local_, off = (uint64_t)comm_->rank() * plane + w0;
                for (int p = 0; p < comm_->n_ranks(); p++)
                    if (peer_val_[p]) dsts.push_back(Engine::GatherDst{peer_val_[p] + off, peer_vld_[p] + off, words_local_});
            }
            if (!dsts.empty() && (rc = e.gather_wires_multi(outputs_.data(), n_out, dsts.data(), (uint32_t)dsts.size()))) return rc;
        }
    }
    for (uint32_t k = 0; k < K_; k++) {
        if ((rc = replicas_[k]->join_copies())) return rc;
        CU_TRY(cudaEventRecord((cudaEvent_t)revents_[k], (cudaStream_t)rstreams_[k]));
        CU_TRY(cudaStreamWaitEvent(ST, (cudaEvent_t)revents_[k], 0));
    }
    if (worst) {
        k_status_max<<<1, 1, 0, ST>>>(d_status_, (uint32_t)worst);
        count_launch();
        CU_TRY(cudaGetLastError());
    }
    worst_ = std::max(worst_, worst);
    last_tiles_ = n_tiles;
    last_tickets_ = 0;
    return B2_OK;
}

int Batch::wait(int *worst) {
    int rc = ensure_device();
    if (rc) return rc;
    if (gather_pending_ && comm_ && comm_->n_ranks() > 1 && (rc = gather_outputs())) return rc;
    CU_TRY(cudaMemcpyAsync(h_status_, d_status_, 4, cudaMemcpyDeviceToHost, ST));
    CU_TRY(cudaStreamSynchronize(ST));
    release_registered();
    if (worst) *worst = std::max<int>(worst_, (int)h_status_[0]);
    worst_ = 0;
    CU_TRY(cudaMemsetAsync(d_status_, 0, 4, ST));
    return B2_OK;
}

}  // namespace b2

// ---------------------------------------------------------------------------- C ABI

#include "capi_types.hpp"

struct b2_batch {
    std::unique_ptr<b2::Batch> b;
};
struct b2_comm {
    std::unique_ptr<b2::Comm> c;
};

#define GUARD_BEGIN try {
#define GUARD_END(err)                        \
    }                                         \
    catch (const std::bad_alloc &) {          \
        return err;                           \
    }                                         \
    catch (...) {                             \
        b2::set_last_error("internal error"); \
        return err;                           \
    }

extern "C" {

int b2_comm_get_unique_id(uint8_t id[128]) {
    if (!id) return B2_ERR_NULL;
    GUARD_BEGIN
    return b2::Comm::unique_id(id);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_comm_init_rank(b2_comm **out, int n_ranks, int rank, const uint8_t id[128], int device) {
    if (!out || !id) return B2_ERR_NULL;
    GUARD_BEGIN
    b2::Comm *c = nullptr;
    int rc = b2::Comm::create(&c, n_ranks, rank, id, device);
    if (rc) return rc;
    *out = new b2_comm();
    (*out)->c.reset(c);
    return B2_OK;
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

void b2_comm_free(b2_comm *c) { delete c; }

int b2_batch_new(const b2_sim *sim, const uint32_t *inputs, uint32_t n_inputs, const uint32_t *outputs, uint32_t n_outputs,
                 uint32_t tile_words, uint32_t n_slots, b2_batch **out) {
    if (!sim || !out || (n_inputs && !inputs) || (n_outputs && !outputs)) return B2_ERR_NULL;
    if (tile_words == 0 || n_slots == 0 || n_slots > 64) return B2_ERR_INVALID_ARG;
    GUARD_BEGIN
    std::unique_ptr<b2::Batch> b(new b2::Batch(*sim->e, inputs, n_inputs, outputs, n_outputs, tile_words, n_slots));
    int rc = b->validate();
    if (rc) return rc;
    *out = new b2_batch();
    (*out)->b = std::move(b);
    return B2_OK;
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

void b2_batch_free(b2_batch *b) { delete b; }
int b2_batch_set_device(b2_batch *b, int device) { return b ? b->b->set_device(device) : B2_ERR_NULL; }
int b2_batch_set_stream(b2_batch *b, void *stream) { return b ? b->b->set_stream(stream) : B2_ERR_NULL; }
int b2_batch_set_option(b2_batch *b, int option, int64_t value) { return b ? b->b->set_option(option, value) : B2_ERR_NULL; }

int b2_batch_get_info(b2_batch *b, b2_batch_info *out) {
    if (!b || !out) return B2_ERR_NULL;
    static_assert(sizeof(b2_batch_info) == sizeof(b2::BatchInfo), "b2_batch_info mirrors b2::BatchInfo");
    return b->b->info((b2::BatchInfo *)out);
}

int b2_batch_run_generated(b2_batch *b, uint64_t seed, int mode, uint64_t lane_word_begin, uint64_t n_words, uint64_t max_steps,
                           uint64_t *out_value, uint64_t *out_valid, size_t out_stride_words, size_t out_ring_words) {
    if (!b) return B2_ERR_NULL;
    if (mode != 0 && mode != 1) return B2_ERR_INVALID_ARG;
    GUARD_BEGIN
    return b->b->run(mode, seed, lane_word_begin, nullptr, nullptr, 0, 0, n_words, max_steps, out_value, out_valid, out_stride_words,
                     out_ring_words);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_batch_run_planes(b2_batch *b, const uint64_t *in_value, const uint64_t *in_valid, size_t in_stride_words, size_t in_ring_words,
                        uint64_t n_words, uint64_t max_steps, uint64_t *out_value, uint64_t *out_valid, size_t out_stride_words,
                        size_t out_ring_words) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b->run(2, 0, 0, in_value, in_valid, in_stride_words, in_ring_words, n_words, max_steps, out_value, out_valid,
                     out_stride_words, out_ring_words);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_batch_wait(b2_batch *b, int *out_worst) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b->wait(out_worst);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_batch_bind_comm(b2_batch *b, b2_comm *comm, uint64_t words_local) {
    if (!b || !comm) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b->bind_comm(comm->c.get(), words_local);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_batch_export_gather_handle(b2_batch *b, uint8_t handle[128]) {
    if (!b || !handle) return B2_ERR_NULL;
    return b->b->export_gather_handle(handle);
}

int b2_batch_import_gather_handles(b2_batch *b, const uint8_t *handles) {
    if (!b || !handles) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b->import_gather_handles(handles);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_batch_gathered(b2_batch *b, uint64_t **dev_value, uint64_t **dev_valid) {
    if (!b || !dev_value || !dev_valid) return B2_ERR_NULL;
    return b->b->gathered(dev_value, dev_valid);
}

int b2_batch_gather_outputs(b2_batch *b) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b->gather_outputs();
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

}  // extern "C"
