// keep_b200 — KEEP forward orchestration: reference call graph (keep_arch.py:1008-1145) re-expressed
// as a stream of hand-written sm_100a kernels over NHWC activations.
#include "engine.h"
#include <cmath>

#include <math.h>
#include <string.h>

#include <algorithm>
#include <mutex>

namespace keep {

// debug launch log (tools/timeline.py): kernel name + grid of every launch, in enqueue order
static std::vector<std::string>* g_launch_log = nullptr;
void launch_log_enable(bool on) {
    delete g_launch_log;
    g_launch_log = on ? new std::vector<std::string>() : nullptr;
}
void launch_log_add(const void* func, dim3 grid, dim3 block) {
    if (!g_launch_log) return;
    const char* nm = nullptr;
    if (cudaFuncGetName(&nm, func) != cudaSuccess || !nm) nm = "?";
    char b[512];
    snprintf(b, sizeof(b), "%s\t%u,%u,%u\t%u", nm, grid.x, grid.y, grid.z, block.x);
    g_launch_log->push_back(b);
}
int launch_log_dump(const char* path) {
    if (!g_launch_log) return -1;
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    for (auto& l : *g_launch_log) fprintf(f, "%s\n", l.c_str());
    fclose(f);
    return (int)g_launch_log->size();
}

// batching of the throughput-bound front half: GMFlow pairs per pass and LQ-encoder frames per pass
static int env_or(const char* name, int dflt, int lo, int hi) {
    const char* e = getenv(name);
    const int v = e ? atoi(e) : dflt;
    return v < lo ? lo : (v > hi ? hi : v);
}
static int flow_chunk() { static const int v = env_or("KEEP_FLOW_CHUNK", 2, 1, 32); return v; }   // measured: 1 / 2 / 3 / 4 pairs -> 165.7 / 165.3 / 162.8 / 162.7 frames/s
static int lq_chunk() { static const int v = env_or("KEEP_LQ_CHUNK", 10, 1, 32); return v; }

bool first_use_on_current_device(unsigned long long* mask) {
    static std::mutex mu;
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    const unsigned long long bit = 1ull << (dev & 63);
    if (*mask & bit) return false;
    *mask |= bit;
    return true;
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("KEEP_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

// =============================================================================================
// Arena
// =============================================================================================
void Arena::reset(char* base, size_t cap, bool dry) {
    base_ = base; cap_ = cap; dry_ = dry; peak_ = 0;
    blocks_.clear();
    blocks_.push_back({0, cap, false});
}

void* Arena::alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    for (size_t i = 0; i < blocks_.size(); ++i) {
        Block& b = blocks_[i];
        if (!b.used && b.size >= bytes) {
            if (b.size > bytes) {
                Block rest{b.off + bytes, b.size - bytes, false};
                b.size = bytes;
                b.used = true;
                const size_t off = b.off;
                blocks_.insert(blocks_.begin() + i + 1, rest);
                peak_ = std::max(peak_, off + bytes);
                return base_ + off;
            }
            b.used = true;
            peak_ = std::max(peak_, b.off + bytes);
            return base_ + b.off;
        }
    }
    throw Error("keep_b200: workspace exhausted (need " + std::to_string(bytes) + " more bytes, capacity " + std::to_string(cap_) +
                "); size it with keep_workspace_bytes(b, T)");
}

void Arena::free(void* p) {
    if (!p) return;
    const size_t off = (size_t)((char*)p - base_);
    for (size_t i = 0; i < blocks_.size(); ++i) {
        if (blocks_[i].off == off && blocks_[i].used) {
            blocks_[i].used = false;
            if (i + 1 < blocks_.size() && !blocks_[i + 1].used) {
                blocks_[i].size += blocks_[i + 1].size;
                blocks_.erase(blocks_.begin() + i + 1);
            }
            if (i > 0 && !blocks_[i - 1].used) {
                blocks_[i - 1].size += blocks_[i].size;
                blocks_.erase(blocks_.begin() + i);
            }
            return;
        }
    }
    throw Error("keep_b200: arena free of unknown pointer");
}

// =============================================================================================
// weights
// =============================================================================================
static bool ends_with(const std::string& s, const std::string& suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

void Engine::add_arr(const std::string& key, const std::vector<float>& host, int d0, int d1, int d2, int d3) {
    staging_.push_back({key, host});
    DevArr a;
    a.numel = (long long)host.size();
    a.d[0] = d0; a.d[1] = d1; a.d[2] = d2; a.d[3] = d3;
    W_[key] = a;
}

// plain conv / linear weights: registered now, uploaded raw and transposed on the device (oihw_to_kc) at the end of pack_weights
void Engine::add_job(const std::string& key, const float* src, int O, int I, int kh, int kw) {
    jobs_.push_back({key, src, O, I, kh * kw});
    DevArr a;
    a.numel = (long long)O * I * kh * kw;
    a.d[0] = I; a.d[1] = O; a.d[2] = kh; a.d[3] = kw;
    W_[key] = a;
}

// conv OIHW -> [(ky*kw + kx)*I + i][o]; linear (O, I) is the 1x1 case  (host version: the few fused / permuted copies)
static std::vector<float> pack_oihw(const float* w, int O, int I, int kh, int kw) {
    std::vector<float> out((size_t)O * I * kh * kw);
    for (int o = 0; o < O; ++o)
        for (int i = 0; i < I; ++i)
            for (int y = 0; y < kh; ++y)
                for (int x = 0; x < kw; ++x)
                    out[((size_t)(y * kw + x) * I + i) * O + o] = w[(((size_t)o * I + i) * kh + y) * kw + x];
    return out;
}

void Engine::pack_weights(const keep_weight_desc* w, int n_w) {
    std::unordered_map<std::string, const keep_weight_desc*> by_name;
    for (int i = 0; i < n_w; ++i) {
        KEEP_CHECK(w[i].name && w[i].data && w[i].ndim >= 1 && w[i].ndim <= 4, "bad weight descriptor %d", i);
        by_name[w[i].name] = &w[i];
    }
    auto numel = [](const keep_weight_desc* d) { long long n = 1; for (int i = 0; i < d->ndim; ++i) n *= d->shape[i]; return n; };
    for (int i = 0; i < n_w; ++i) {
        const keep_weight_desc* d = &w[i];
        const std::string key = d->name;
        const long long ne = numel(d);
        if (key == "position_emb" || key == "quantize.embedding.weight") {
            add_arr(key, std::vector<float>(d->data, d->data + ne), (int)d->shape[0], (int)d->shape[1]);
        } else if (ends_with(key, "in_proj_weight")) {  // nn.MultiheadAttention packed (3E, E): q | k | v
            const int E = (int)d->shape[1];
            KEEP_CHECK(d->ndim == 2 && d->shape[0] == 3 * E, "in_proj_weight shape");
            const std::string base = key.substr(0, key.size() - strlen("in_proj_weight"));
            add_arr(base + "in_proj_qk.weight", pack_oihw(d->data, 2 * E, E, 1, 1), E, 2 * E, 1, 1);
            add_arr(base + "in_proj_v.weight", pack_oihw(d->data + (size_t)2 * E * E, E, E, 1, 1), E, E, 1, 1);
        } else if (ends_with(key, "in_proj_bias")) {
            const int E = (int)(d->shape[0] / 3);
            const std::string base = key.substr(0, key.size() - strlen("in_proj_bias"));
            add_arr(base + "in_proj_qk.bias", std::vector<float>(d->data, d->data + 2 * E), 2 * E);
            add_arr(base + "in_proj_v.bias", std::vector<float>(d->data + 2 * E, d->data + 3 * E), E);
        } else if (d->ndim == 4) {
            add_job(key, d->data, (int)d->shape[0], (int)d->shape[1], (int)d->shape[2], (int)d->shape[3]);
            // GMFlow upsampler conv on cat(flow[2], feature[128]) (gmflow/gmflow.py:46-48,76): a tensor-core copy with the
            // input channels reordered to [feature 128 | flow 2 | 6 zeros] = 136, so both sources start on 8-channel units
            if (ends_with(key, ".upsampler.0.weight") && d->shape[1] == 130 && d->shape[2] == 3) {
                const int O = (int)d->shape[0];
                std::vector<float> perm((size_t)O * 136 * 9, 0.0f);
                for (int o = 0; o < O; ++o)
                    for (int i = 0; i < 130; ++i) {
                        const int ni = i < 2 ? 128 + i : i - 2;
                        for (int t = 0; t < 9; ++t) perm[((size_t)o * 136 + ni) * 9 + t] = d->data[((size_t)o * 130 + i) * 9 + t];
                    }
                add_arr(key.substr(0, key.size() - strlen("weight")) + "perm136.weight", pack_oihw(perm.data(), O, 136, 3, 3), 136, O, 3, 3);
            }
            // AttnBlock q|k|v 1x1 convs fused into one N = 3C GEMM (vqgan_arch.py:222-224)
            if (ends_with(key, ".q.weight")) {
                const std::string base = key.substr(0, key.size() - strlen("q.weight"));
                auto kq = by_name.find(base + "k.weight"), vq = by_name.find(base + "v.weight");
                auto bq = by_name.find(base + "q.bias"), bk = by_name.find(base + "k.bias"), bv = by_name.find(base + "v.bias");
                if (kq != by_name.end() && vq != by_name.end() && bq != by_name.end() && bk != by_name.end() && bv != by_name.end()) {
                    const int C = (int)d->shape[1], O = (int)d->shape[0];
                    std::vector<float> cat((size_t)C * 3 * O), bias(3 * O);
                    const float* src[3] = {d->data, kq->second->data, vq->second->data};
                    const float* bsrc[3] = {bq->second->data, bk->second->data, bv->second->data};
                    for (int t = 0; t < 3; ++t) {
                        for (int o = 0; o < O; ++o) {
                            for (int c = 0; c < C; ++c) cat[(size_t)c * 3 * O + t * O + o] = src[t][(size_t)o * C + c];
                            bias[t * O + o] = bsrc[t][o];
                        }
                    }
                    add_arr(base + "qkv.weight", cat, C, 3 * O, 1, 1);
                    add_arr(base + "qkv.bias", bias, 3 * O);
                }
            }
        } else if (d->ndim == 2) {
            add_job(key, d->data, (int)d->shape[0], (int)d->shape[1], 1, 1);
            // GMFlow attention projections (gmflow/transformer.py:117-119, bias-free): fused copies q|k|v (self-attention:
            // one source) and k|v (cross-attention: both read the target) -> one N = 3C / 2C GEMM that converts the
            // activations once (A-stationary walk over the N tiles) instead of three / two times (KEEP_GM_FUSE_QKV=1)
            if (ends_with(key, ".q_proj.weight") && key.find("flownet.") == 0) {
                const std::string base = key.substr(0, key.size() - strlen("q_proj.weight"));
                auto kq = by_name.find(base + "k_proj.weight"), vq = by_name.find(base + "v_proj.weight");
                if (kq != by_name.end() && vq != by_name.end() && kq->second->ndim == 2 && vq->second->ndim == 2 &&
                    kq->second->shape[0] == d->shape[0] && vq->second->shape[0] == d->shape[0] &&
                    kq->second->shape[1] == d->shape[1] && vq->second->shape[1] == d->shape[1]) {
                    const int C = (int)d->shape[1], O = (int)d->shape[0];
                    std::vector<float> qkv((size_t)C * 3 * O), kv((size_t)C * 2 * O);
                    const float* src[3] = {d->data, kq->second->data, vq->second->data};
                    for (int t = 0; t < 3; ++t)
                        for (int o = 0; o < O; ++o)
                            for (int c = 0; c < C; ++c) {
                                qkv[(size_t)c * 3 * O + t * O + o] = src[t][(size_t)o * C + c];
                                if (t > 0) kv[(size_t)c * 2 * O + (t - 1) * O + o] = src[t][(size_t)o * C + c];
                            }
                    add_arr(base + "qkv_proj.weight", qkv, C, 3 * O, 1, 1);
                    add_arr(base + "kv_proj.weight", kv, C, 2 * O, 1, 1);
                }
            }
        } else {
            add_arr(key, std::vector<float>(d->data, d->data + ne), (int)d->shape[0]);
        }
    }
    // one pool.  Small host-packed arrays are copied as they are; the bulk (jobs_) goes up raw into a temporary buffer and is
    // transposed by a device kernel -- the host never touches the 158 M parameters element by element
    size_t total = 0, raw_total = 0;
    for (auto& kv : staging_) total += (kv.second.size() + 63) & ~(size_t)63;
    for (auto& j : jobs_) { const size_t ne = (size_t)j.O * j.I * j.taps; total += (ne + 63) & ~(size_t)63; raw_total += (ne + 63) & ~(size_t)63; }
    float* raw = nullptr;
    if (!dry_only_) {
        CUDA_CHECK(cudaMalloc((void**)&wpool_, total * sizeof(float)));
        CUDA_CHECK(cudaMalloc((void**)&raw, std::max<size_t>(raw_total, 64) * sizeof(float)));
    }
    size_t off = 0, roff = 0;
    for (auto& kv : staging_) {
        if (!dry_only_)
            CUDA_CHECK(cudaMemcpyAsync(wpool_ + off, kv.second.data(), kv.second.size() * sizeof(float), cudaMemcpyHostToDevice, 0));
        W_[kv.first].p = (dry_only_ ? (float*)(uintptr_t)4096 : wpool_) + off;
        off += (kv.second.size() + 63) & ~(size_t)63;
    }
    for (auto& j : jobs_) {
        const size_t ne = (size_t)j.O * j.I * j.taps;
        if (!dry_only_) {
            CUDA_CHECK(cudaMemcpyAsync(raw + roff, j.src, ne * sizeof(float), cudaMemcpyHostToDevice, 0));
            if (j.taps == 1 && (j.O == 1 || j.I == 1)) CUDA_CHECK(cudaMemcpyAsync(wpool_ + off, raw + roff, ne * sizeof(float), cudaMemcpyDeviceToDevice, 0));
            else oihw_to_kc(raw + roff, wpool_ + off, j.O, j.I, j.taps, 0);
        }
        W_[j.key].p = (dry_only_ ? (float*)(uintptr_t)4096 : wpool_) + off;
        off += (ne + 63) & ~(size_t)63;
        roff += (ne + 63) & ~(size_t)63;
    }
    if (!dry_only_) {
        CUDA_CHECK(cudaStreamSynchronize(0));   // host staging vectors and the caller's tensors are read until here
        CUDA_CHECK(cudaGetLastError());
        cudaFree(raw);
    }
    staging_.clear();
    staging_.shrink_to_fit();
    jobs_.clear();
}

const float* Engine::warr(const std::string& key) const {
    auto it = W_.find(key);
    KEEP_CHECK(it != W_.end(), "missing weight '%s' (state dict must match the reference's KEEP keys)", key.c_str());
    return it->second.p;
}

ConvW Engine::convw(const std::string& prefix) const {
    auto it = W_.find(prefix + ".weight");
    KEEP_CHECK(it != W_.end(), "missing weight '%s.weight'", prefix.c_str());
    ConvW c;
    c.w = it->second.p;
    c.cin = it->second.d[0]; c.cout = it->second.d[1]; c.kh = it->second.d[2]; c.kw = it->second.d[3];
    KEEP_CHECK(c.kh >= 1 && c.kw >= 1, "'%s.weight' is not a conv/linear weight", prefix.c_str());
    auto ib = W_.find(prefix + ".bias");
    c.b = ib == W_.end() ? nullptr : ib->second.p;
    return c;
}

// fusion points: feature size -> channels (keep_arch.py:940-947), encoder block whose output is tapped (:950-951) and
// generator block after which CFT / CFA run (:953-954)
static const int kFuseSize[6] = {16, 32, 64, 128, 256, 512};
static const int kFuseCh[6] = {512, 256, 256, 128, 128, 64};
static const int kFuseEnc[6] = {18, 14, 11, 8, 5, 2};
static const int kFuseGen[6] = {6, 9, 12, 15, 18, 21};

// =============================================================================================
// construction
// =============================================================================================
Engine::Engine(int device, const keep_weight_desc* w, int n_w, int flags) : device_(device), flags_(flags) {
    // KEEP_FORCE_FLAGS=<int>: OR extra engine flags in (A/B runs on the GPU box, e.g. 16 = KEEP_FLAG_TC_WIDE on the general config)
    if (const char* e = getenv("KEEP_FORCE_FLAGS")) { flags |= atoi(e) & ~KEEP_FLAG_PLAN_ONLY; flags_ = flags; }
    dry_only_ = (flags & KEEP_FLAG_PLAN_ONLY) != 0;
    // fp16 feature-map storage is wired through the kernels but not through the whole programme (CFA and the cross-frame
    // state are fp32-only) and it cannot meet the parity bar with fp32-grade operands anyway: refuse it loudly
    KEEP_CHECK(!(flags & KEEP_FLAG_FP16_FEATURES), "KEEP_FLAG_FP16_FEATURES is not supported by this engine (feature maps are fp32)");
    if (!dry_only_) {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        KEEP_CHECK(e == cudaSuccess && ndev > 0, "keep_b200 needs a CUDA device (no CPU fallback): %s", cudaGetErrorString(e));
        KEEP_CHECK(device >= 0 && device < ndev, "device %d out of range (have %d)", device, ndev);
        CUDA_CHECK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        KEEP_CHECK(prop.major == 10, "keep_b200 is built for sm_100a (Blackwell B200); device %d is sm_%d%d", device,
                   prop.major, prop.minor);
        num_sms_ = prop.multiProcessorCount;
        main_cap_ = num_sms_;
        gn_warmup();
        CUDA_CHECK(cudaMalloc((void**)&gn_tickets_, 2 * gn_ticket_count() * sizeof(int)));   // main / side stream
        CUDA_CHECK(cudaMemset(gn_tickets_, 0, 2 * gn_ticket_count() * sizeof(int)));
        CUDA_CHECK(cudaMalloc((void**)&status_, sizeof(int)));
        CUDA_CHECK(cudaMemset(status_, 0, sizeof(int)));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_last_, cudaEventDisableTiming));
        // KEEP_SIDE_SMS = n > 0: side-branch persistent kernels capped at n CTAs; n < 0: short CTAs of -n work items each
        { const char* e = getenv("KEEP_SIDE_SMS"); side_sms_ = e ? atoi(e) : 64; if (side_sms_ > num_sms_ || (side_sms_ >= 0 && side_sms_ < 8)) side_sms_ = num_sms_; }
    }
    { const char* e = getenv("KEEP_BATCH_MAX"); batch_max_ = e ? std::max(1, std::min(8, atoi(e))) : 2; }
    adt_ = (flags & KEEP_FLAG_FP16_FEATURES) ? F16 : F32;
    tc_passes_ = (flags & KEEP_FLAG_TC_SPLIT3) ? 3 : 1;
    pack_weights(w, n_w);
    // required keys (strict load, like load_state_dict(strict=True))
    const char* must[] = {"position_emb", "quantize.embedding.weight", "feat_emb.weight", "idx_pred_layer.1.weight",
                          "encoder.blocks.0.weight", "hq_encoder.blocks.24.weight", "generator.blocks.24.weight",
                          "flownet.model.backbone.conv1.weight", "kalman_filter.kalman_gain_calculator.3.weight",
                          "cfa.16.attn.to_q.weight", "cfa.32.attn.to_q.weight", "ft_layers.8.linear2.weight"};
    for (const char* k : must) KEEP_CHECK(has(k), "state dict is missing '%s'", k);
    // fusion points are read off the tensor names, like the reference builds cft / cfa from cft_list / cfa_list
    // (keep_arch.py:957-967): a size is fused iff its block is in the state dict, and then the block must be complete
    int n_cft = 0;
    for (int i = 0; i < 6; ++i) {
        const std::string sz = std::to_string(kFuseSize[i]);
        cft_on_[i] = has("cft." + sz + ".scale.0.weight") || has("cft." + sz + ".encode_enc.conv1.weight");
        cfa_on_[i] = has("cfa." + sz + ".attn.to_q.weight");
        if (cft_on_[i]) {
            ++n_cft;
            for (const char* k : {".encode_enc.norm1.weight", ".encode_enc.conv1.weight", ".encode_enc.norm2.weight",
                                  ".encode_enc.conv2.weight", ".encode_enc.conv_out.weight", ".scale.0.weight", ".scale.2.weight",
                                  ".shift.0.weight", ".shift.2.weight"})
                KEEP_CHECK(has("cft." + sz + k), "state dict is missing 'cft.%s%s'", sz.c_str(), k);
            const DevArr& w = W_.at("cft." + sz + ".scale.0.weight");
            KEEP_CHECK(w.d[0] == kFuseCh[i] && w.d[1] == kFuseCh[i], "cft.%s.scale.0.weight: expected %d channels, got (%d, %d)",
                       sz.c_str(), kFuseCh[i], w.d[0], w.d[1]);
        }
    }
    KEEP_CHECK(n_cft > 0, "state dict has no 'cft.<size>.*' block (KEEP fuses at 16/32/64, Asian at 32/64/128/256)");
    KEEP_CHECK(!cfa_on_[2] && !cfa_on_[3] && !cfa_on_[4] && !cfa_on_[5], "cfa blocks beyond 16/32 are not part of either reference config");

    // GMFlow shifted-window region ids (gmflow/transformer.py:19-43) for a 64x64 map, 2x2 windows, shift 16
    if (!dry_only_) {
        const int H = 64, Wd = 64, wh = 32, ww = 32, sh = 16, sw = 16;
        std::vector<int> reg(4 * 1024);
        for (int win = 0; win < 4; ++win)
            for (int t = 0; t < 1024; ++t) {
                const int y = (win / 2) * wh + t / ww, x = (win % 2) * ww + t % ww;
                const int ry = y < H - wh ? 0 : (y < H - sh ? 1 : 2);
                const int rx = x < Wd - ww ? 0 : (x < Wd - sw ? 1 : 2);
                reg[win * 1024 + t] = ry * 3 + rx;
            }
        CUDA_CHECK(cudaMalloc((void**)&region_, reg.size() * sizeof(int)));
        CUDA_CHECK(cudaMemcpy(region_, reg.data(), reg.size() * sizeof(int), cudaMemcpyHostToDevice));
        std::vector<unsigned char> reg8(reg.begin(), reg.end());
        CUDA_CHECK(cudaMalloc((void**)&region8_, reg8.size()));
        CUDA_CHECK(cudaMemcpy(region8_, reg8.data(), reg8.size(), cudaMemcpyHostToDevice));
        std::vector<float> grid(4096 * 2);
        for (int t = 0; t < 4096; ++t) { grid[2 * t] = (float)(t % 64); grid[2 * t + 1] = (float)(t / 64); }
        CUDA_CHECK(cudaMalloc((void**)&grid64_, grid.size() * sizeof(float)));
        CUDA_CHECK(cudaMemcpy(grid64_, grid.data(), grid.size() * sizeof(float), cudaMemcpyHostToDevice));
        // streams / events a forward needs: created here so that nothing is created while a graph is being captured
        int lo = 0, hi = 0;
        CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));       // lo = least urgent
        CUDA_CHECK(cudaStreamCreateWithPriority(&side_, cudaStreamNonBlocking, lo));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
        CUDA_CHECK(cudaStreamCreateWithPriority(&gs_, cudaStreamNonBlocking, hi));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_in_, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_out_, cudaEventDisableTiming));
        for (int i = 0; i < 64; ++i) {
            cudaEvent_t e;
            CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ev_flow_.push_back(e);
        }
        tc_configure_device();
        conv_small_configure_device();
        attention_tc_configure_device();
        if ((flags_ & KEEP_FLAG_TCGEN05) && tc_passes_ == 3) prepack_tc_weights();
    }
}

// tcgen05 weight panels for every layer of a forward, packed once at creation: a dry run (no device work) of a 2-frame clip
// lists the (layer, N tile, passes, stride-2 pad, wide) variants -- in the split-precision mode the N tile depends on the
// layer only, not on the batch, so the list holds for every T and for lockstep groups
void Engine::prepack_tc_weights() {
    std::vector<TcReq> reqs;
    tc_collect_ = &reqs;
    try {
        begin(nullptr, 0, nullptr, true);
        forward_clip(nullptr, 2, nullptr, KEEP_OUT_F32);
    } catch (...) {
        tc_collect_ = nullptr;
        throw;
    }
    tc_collect_ = nullptr;
    std::vector<TcReq> uniq;
    std::vector<size_t> offs;
    size_t total = 0;
    for (const TcReq& r : reqs) {
        bool seen = false;
        for (const TcReq& u : uniq) seen = seen || (u.cw.w == r.cw.w && u.bn == r.bn && u.passes == r.passes && u.wide == r.wide);
        if (seen) continue;
        const int cb = tc_cb(r.passes);
        const int vcin = r.s2d_pad >= 0 ? 4 * ((r.cw.cin + cb - 1) / cb) * cb : r.cw.cin;
        const int vtaps = r.s2d_pad >= 0 ? 4 : r.cw.kh * r.cw.kw;
        uniq.push_back(r);
        offs.push_back(total);
        total += (tc_packed_weight_halfs(vcin, r.cw.cout, vtaps, r.bn, r.passes) + 127) & ~(size_t)127;
    }
    if (uniq.empty()) return;
    CUDA_CHECK(cudaMalloc((void**)&tcw_pool_, total * sizeof(__half)));
    for (size_t i = 0; i < uniq.size(); ++i) {
        const TcReq& r = uniq[i];
        TcW t;
        t.bn = r.bn; t.passes = r.passes; t.wide = r.wide; t.pooled = true;
        t.p = tcw_pool_ + offs[i];
        tc_repack_device(r.cw.w, r.cw.cin, r.cw.cout, r.cw.kh * r.cw.kw, r.bn, r.passes, r.s2d_pad, t.p, 0, r.wide);
        tcw_[r.cw.w].push_back(t);
    }
    CUDA_CHECK(cudaStreamSynchronize(0));
    prepacked_ = true;
}

Engine::~Engine() {
    if (dry_only_) return;
    cudaSetDevice(device_);
    cudaFree(wpool_);
    cudaFree(region_);
    cudaFree(region8_);
    cudaFree(gn_tickets_);
    cudaFree(status_);
    if (ev_last_) cudaEventDestroy(ev_last_);
    cudaFree(u8_stage_);
    cudaFree(grid64_);
    cudaFree(own_ws_);
    for (auto& kv : cap_) cudaFree(kv.second.p);
    for (auto& kv : forced_) cudaFree(kv.second.p);
    for (auto& kv : tcw_)
        for (auto& v : kv.second)
            if (!v.pooled) cudaFree(v.p);
    cudaFree(tcw_pool_);
    for (auto& g : graphs_) cudaGraphExecDestroy(g.second.exec);
    cudaFree(gx_);
    cudaFree(gout_);
    if (gs_) { cudaStreamDestroy(gs_); cudaEventDestroy(ev_in_); cudaEventDestroy(ev_out_); }
    if (side_) { cudaStreamDestroy(side_); cudaEventDestroy(ev_fork_); }
    for (auto& e : ev_flow_) cudaEventDestroy(e);
    for (auto& p : prof_) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto& e : ev_pool_) cudaEventDestroy(e);
}

// =============================================================================================
// building blocks
// =============================================================================================
// The workspace is split in two arenas: the main one and a side one for the GMFlow branch, which runs on its own
// stream concurrently with the serial per-frame chain (stream-ordered reuse is only safe within one stream).
void Engine::begin(void* ws, size_t ws_bytes, cudaStream_t s, bool dry) {
    s_ = s_main_ = s;
    ar_ = &arena_;
    if (dry) {
        arena_.reset((char*)(uintptr_t)4096, (size_t)1 << 45, true);
        arena2_.reset((char*)(uintptr_t)4096 + ((size_t)1 << 45), (size_t)1 << 45, true);
    } else {
        KEEP_CHECK(side_bytes_ > 0 && side_bytes_ < ws_bytes, "workspace split not planned");
        arena_.reset((char*)ws, ws_bytes - side_bytes_, false);
        arena2_.reset((char*)ws + (ws_bytes - side_bytes_), side_bytes_, false);
    }
}

Tensor Engine::talloc(int n, int h, int w, int c, int dt) {
    Tensor t;
    t.n = n; t.h = h; t.w = w; t.c = c; t.dt = (DType)dt;
    t.p = ar_->alloc(t.bytes());
    return t;
}
void Engine::tfree(Tensor& t) {
    if (t.gn_part) { ar_->free(t.gn_part); t.gn_part = nullptr; }
    if (t.aff) { ar_->free(t.aff); t.aff = nullptr; }
    ar_->free(t.p);
    t.p = nullptr;
}
void Engine::afree(Aff& a) { ar_->free(a.scale); a.scale = a.shift = nullptr; }

void Engine::capture(const std::string& name, const void* dev, size_t bytes) {
    if (!capture_ || ar_->dry()) return;
    Cap& c = cap_[name];
    if (c.bytes < bytes) {
        cudaFree(c.p);
        CUDA_CHECK(cudaMalloc(&c.p, bytes));
        c.bytes = bytes;
    }
    CUDA_CHECK(cudaMemcpyAsync(c.p, dev, bytes, cudaMemcpyDeviceToDevice, s_));
}

Tensor Engine::conv(const Tensor& x, const ConvW& cw, const ConvOpt& o) {
    const int c1 = o.in1 ? o.in1->c : 0;
    KEEP_CHECK(cw.cin == x.c + c1, "conv: weight expects %d input channels, got %d", cw.cin, x.c + c1);
    ConvArgs a;
    a.in0 = x.p; a.in0_dt = x.dt; a.c0 = x.c;
    if (o.in1) {
        KEEP_CHECK(o.in1->n == x.n && o.in1->h == x.h && o.in1->w == x.w, "conv: concat sources differ in shape");
        a.in1 = o.in1->p; a.in1_dt = o.in1->dt; a.c1 = c1;
    }
    a.n = x.n; a.h = x.h; a.w = x.w; a.up = o.up;
    if (o.pre) { a.pre_scale = o.pre->scale; a.pre_shift = o.pre->shift; }
    a.pre_act = o.pre_act;
    a.wt = cw.w; a.bias = cw.b;
    a.kh = cw.kh; a.kw = cw.kw; a.stride = o.stride; a.pad_t = o.pad_t; a.pad_l = o.pad_l; a.cout = cw.cout;
    a.ho = (x.h * o.up + o.pad_t + o.pad_b - cw.kh) / o.stride + 1;
    a.wo = (x.w * o.up + o.pad_l + o.pad_r - cw.kw) / o.stride + 1;
    a.act = o.act;
    Tensor out = talloc(x.n, a.ho, a.wo, cw.cout, o.out_dt < 0 ? adt_ : o.out_dt);
    if (o.res) {
        KEEP_CHECK(o.res->numel() == out.numel(), "conv: residual shape mismatch");
        a.res = o.res->p; a.res_dt = o.res->dt;
    }
    a.out = out.p; a.out_dt = out.dt;
    const bool use_small = conv_small_eligible(a);
    const bool use_tc = !use_small && (flags_ & KEEP_FLAG_TCGEN05) && !o.exact && tc_eligible(a);
    int bn = 0;
    void* part = nullptr;
    // precision of this layer's tensor-core operands: the engine mode, or 1 pass where the caller marked the layer as
    // downstream of every decision (pass_override_, see generator())
    const int passes = (pass_override_ && tc_passes_ == 3) ? pass_override_ : tc_passes_;
    a.pre_exact = (passes == 1 && tc_passes_ == 3) ? 1 : 0;   // keep the exact fp32 swish in front of the fp16 rounding
    // un-normalised feature maps in the generator (upsample convs, 1x1 shortcuts, CFT scale / shift convs) have no bound on
    // their magnitude: four stacked CFT modulations of the 'Asian' programme reach 6e4 with synthetic weights, past fp16
    a.a_wide = (wide_scope_ && !o.pre && x.w > 1) ? 1 : 0;
    if (use_tc) {
        const long long m_tiles = cw.kh == 3 ? (long long)x.n * cdiv(a.ho, 16) * cdiv(a.wo, 8)
                                             : (long long)x.n * cdiv((long long)x.h * x.w, 128);
        bn = tc_pick_bn(cw.cout, m_tiles, passes);
        a.splitk = tc_pick_splitk(m_tiles, cdiv(cw.cout, bn), cdiv(tc_virtual_cin(a, passes), tc_cb(passes)));
    } else {
        a.splitk = use_small ? 1 : conv_pick_splitk(a);
    }
    if (a.splitk > 1) {
        part = ar_->alloc((size_t)a.splitk * out.numel() * sizeof(float));
        a.partial = (float*)part;
    }
    // GroupNorm statistics of the output from the producing kernel (conv epilogue, or the split-K reduce): the consumer's
    // gn() then needs one tiny finalize launch instead of a read pass over the tensor plus a second kernel
    // KEEP_GN_EPILOGUE: 0 = stand-alone statistics kernels everywhere; 1 (default) = every eligible layer; 2 = split-K layers
    // only (the reduce kernel emits them almost for free); 3 = split-K layers + conv epilogues with >= 128 output channels.
    // Measured on B200 (same box, frames/s at T = 20): 146.2 / 151.0 / 149.1 / 149.3 -- the N = 64 layers pay ~18 us per launch
    // for the extra ~100 epilogue instructions per 16-column chunk (their epilogue is on the critical path), the read pass
    // and second kernel they replace cost more.
    static const int gn_mode = getenv("KEEP_GN_EPILOGUE") ? atoi(getenv("KEEP_GN_EPILOGUE")) : 1;
    static const bool cluster_mode = getenv("KEEP_TC_CLUSTER") && atoi(getenv("KEEP_TC_CLUSTER")) >= 2;   // (no reduce kernel to emit them)
    const bool gn_here = gn_mode == 1 || (gn_mode >= 2 && a.splitk > 1) || (gn_mode == 3 && cw.cout >= 128);
    if (o.want_stats && use_tc && gn_here && !(cluster_mode && a.splitk > 1)) {
        const int P = conv_gn_slots(a, a.splitk);
        if (P > 0) {
            out.gn_P = P;
            out.gn_part = (float*)ar_->alloc((size_t)x.n * P * 64 * sizeof(float));
            a.gn_part = out.gn_part; a.gn_P = P;
            // KEEP_GN_REDUCE_FINAL=1 (opt-in): few slots behind a split-K reduce -> its last block per image finalizes them (ticket), no
            // gn_finalize_parts launch.  Measured 0.6 % SLOWER than the separate launch (176.9 vs 177.9 frames/s): under graph + PDL
            // the launch boundary costs less than the serial tail (fence, ticket, one block walking 32 x P slots)
            static const bool fin_en = getenv("KEEP_GN_REDUCE_FINAL") && getenv("KEEP_GN_REDUCE_FINAL")[0] == '1';
            if (fin_en && a.splitk > 1 && P <= kGnReduceFinalMaxP && !o.stats_norm.empty() && x.n <= gn_ticket_count() &&
                has(o.stats_norm + ".weight")) {
                out.aff = (float*)ar_->alloc((size_t)2 * x.n * cw.cout * sizeof(float));
                out.aff_gamma = warr(o.stats_norm + ".weight");
                a.gn_fin_gamma = out.aff_gamma; a.gn_fin_beta = warr(o.stats_norm + ".bias");
                a.gn_fin_scale = out.aff; a.gn_fin_shift = out.aff + (size_t)x.n * cw.cout;
                a.gn_tickets = gn_tickets_ ? gn_tickets_ + ((s_ == side_ && side_) ? gn_ticket_count() : 0) : nullptr;
            }
        }
    }
    // LayerNorm of the output rows inside the split-K reduce (transformer linears; KEEP_LN_REDUCE=0: separate layernorm launch)
    static const bool ln_en = !(getenv("KEEP_LN_REDUCE") && getenv("KEEP_LN_REDUCE")[0] == '0');
    if (o.ln && ln_en && use_tc && a.splitk > 1 && !cluster_mode && out.dt == F32 && !a.gn_part &&
        splitk_reduce_ln_eligible((long long)out.numel(), cw.cout)) {
        LnFuse& f = *o.ln;
        f.out = talloc(out.n, out.h, out.w, out.c, F32);
        if (f.add2) f.out2 = talloc(out.n, out.h, out.w, out.c, F32);
        a.ln_g = warr(f.prefix + ".weight"); a.ln_b = warr(f.prefix + ".bias"); a.ln_eps = 1e-5f; a.ln_out = f.out.f();
        a.ln_add2 = f.add2; a.ln_add2_rows = f.add2_rows; a.ln_out2 = f.add2 ? f.out2.f() : nullptr;
        f.done = true;
        if (plan_) plan_->push_back("layernorm rows=" + std::to_string(out.rows()) + " c=" + std::to_string(out.c) + " fused=1");
    }
    if (tc_collect_ && use_tc)
        tc_collect_->push_back({cw, bn, passes, tc_is_s2d(a) ? a.pad_t : -1, (wide_scope_ && !o.pre && x.w > 1 && passes == 3 && a.in0_dt == F32) ? 1 : 0});
    if (plan_) {
        char line[256];
        snprintf(line, sizeof(line), "conv n=%d h=%d w=%d c0=%d c1=%d cout=%d k=%d stride=%d up=%d pre=%d act=%d res=%d kernel=%s splitk=%d bn=%d wide=%d",
                 a.n, a.h, a.w, a.c0, a.c1, a.cout, a.kh, a.stride, a.up, a.pre_scale ? a.pre_act + 1 : 0, a.act, a.res ? 1 : 0,
                 use_tc ? "tcgen05" : (use_small ? "small" : "simt"), a.splitk, bn, (a.a_wide && use_tc && passes == 3) ? 1 : 0);
        plan_->push_back(std::string(line) + (a.gn_fin_scale ? " gnstats=2" : (a.gn_part ? " gnstats=1" : "")));
    }
    if (!ar_->dry()) {
        Prof pr;
        if (profile_) {
            pr.a = get_event(); pr.b = get_event();
            pr.flops = 2.0 * (double)out.rows() * cw.cout * cw.kh * cw.kw * cw.cin;
            pr.bytes = (double)x.bytes() + (o.in1 ? (double)o.in1->bytes() : 0.0) + (double)out.bytes() +
                       (o.res ? (double)o.res->bytes() : 0.0) + (use_tc ? 2.0 * (passes == 3 ? 2 : 1) : 4.0) * cw.cout * cw.kh * cw.kw * cw.cin;
            pr.tag = use_tc ? 1 : 0;
            pr.m = (int)out.rows(); pr.k = cw.kh * cw.kw * cw.cin; pr.n = cw.cout; pr.kh = cw.kh * 10 + o.stride; pr.splitk = a.splitk; pr.bn = bn;
            CUDA_CHECK(cudaEventRecord(pr.a, s_));
        }
        // side-branch (GMFlow) kernels are persistent too: cap their grid so the latency-critical serial chain on the main
        // stream always finds free SMs
        const int grid_cap = (s_ == side_ && side_) ? side_sms_ : main_cap_;
        int nl = a.splitk > 1 ? 2 : 1;
        if (use_tc) {
            a.a_wide = (a.a_wide && passes == 3 && a.in0_dt == F32) ? 1 : 0;
            bool pooled = false;
            const __half* panels = tc_weights(cw, bn, passes, tc_is_s2d(a) ? a.pad_t : -1, a.a_wide, &pooled);
            a.wt_static = pooled ? 1 : 0;   // packed at engine creation: the loader may prefetch ahead of griddepcontrol.wait
            nl = conv2d_tc(a, panels, bn, passes, a.splitk, a.partial, grid_cap, s_);
        }
        else if (use_small) conv2d_small(a, s_);
        else conv2d_simt(a, s_);
        launches_ += nl;
        if (profile_) {
            CUDA_CHECK(cudaEventRecord(pr.b, s_));
            prof_.push_back(pr);
        }
        // KEEP_DEBUG_VERIFY_TC=1 (debug, eager mode only): re-run every tcgen05 layer on the exact-fp32 CUDA-core kernel and
        // report the layers whose results differ -- pinpoints a bad (shape, tile, split) configuration in one forward
        static const bool verify = getenv("KEEP_DEBUG_VERIFY_TC") != nullptr;
        if (verify && use_tc && out.dt == F32) {
            CUDA_CHECK(cudaStreamSynchronize(s_));
            ConvArgs b = a;
            float* ref = nullptr;
            CUDA_CHECK(cudaMalloc((void**)&ref, out.bytes()));
            b.out = ref; b.splitk = 1; b.partial = nullptr;
            conv2d_simt(b, s_);
            CUDA_CHECK(cudaStreamSynchronize(s_));
            std::vector<float> h_tc(out.numel()), h_ref(out.numel());
            CUDA_CHECK(cudaMemcpy(h_tc.data(), out.p, out.bytes(), cudaMemcpyDeviceToHost));
            CUDA_CHECK(cudaMemcpy(h_ref.data(), ref, out.bytes(), cudaMemcpyDeviceToHost));
            cudaFree(ref);
            double emax = 0.0, rmax = 0.0;
            long long bad = 0, first_bad = -1;
            for (size_t i = 0; i < h_tc.size(); ++i) {
                if (!std::isfinite(h_tc[i])) { if (bad++ == 0) first_bad = (long long)i; continue; }
                emax = std::max(emax, (double)std::fabs(h_tc[i] - h_ref[i]));
                rmax = std::max(rmax, (double)std::fabs(h_ref[i]));
            }
            static long long checked = 0, flagged = 0;
            ++checked;
            const bool flag = bad > 0 || emax > 2e-3 * std::max(1.0, rmax);
            if (flag) ++flagged;
            if (flag || checked % 200 == 0)
                fprintf(stderr, "[verify_tc] %s layer %lld: n=%d h=%d w=%d c0=%d c1=%d cout=%d k=%d stride=%d up=%d pre=%d/%d act=%d res=%d "
                        "splitk=%d bn=%d passes=%d | max|err|=%.3g max|ref|=%.3g nonfinite=%lld (first at %lld) [flagged %lld of %lld]\n",
                        flag ? "MISMATCH" : "ok", checked, a.n, a.h, a.w, a.c0, a.c1, a.cout, a.kh, a.stride, a.up, a.pre_scale ? 1 : 0,
                        a.pre_act, a.act, a.res ? 1 : 0, a.splitk, bn, passes, emax, rmax, bad, first_bad, flagged, checked);
        }
    }
    if (part) ar_->free(part);
    return out;
}

const __half* Engine::tc_weights(const ConvW& cw, int bn, int passes, int s2d_pad, int wide, bool* pooled) {
    std::vector<TcW>& variants = tcw_[cw.w];
    if (pooled) *pooled = false;
    for (const TcW& v : variants)
        if (v.bn == bn && v.passes == passes && v.wide == wide) {
            if (pooled) *pooled = v.pooled;
            return v.p;
        }
    TcW t;
    t.bn = bn; t.passes = passes; t.wide = wide;
    const int cb = tc_cb(passes);
    const int vcin = s2d_pad >= 0 ? 4 * ((cw.cin + cb - 1) / cb) * cb : cw.cin;
    const int vtaps = s2d_pad >= 0 ? 4 : cw.kh * cw.kw;
    CUDA_CHECK(cudaMalloc((void**)&t.p, tc_packed_weight_halfs(vcin, cw.cout, vtaps, bn, passes) * sizeof(__half)));
    tc_repack_device(cw.w, cw.cin, cw.cout, cw.kh * cw.kw, bn, passes, s2d_pad, t.p, s_, wide);
    variants.push_back(t);
    return t.p;
}

Tensor Engine::linear(const Tensor& x, const std::string& prefix, int act, const Tensor* res, LnFuse* lnf) {
    ConvOpt o;
    o.act = act; o.res = res; o.out_dt = F32; o.ln = lnf;
    return conv(x, convw(prefix), o);
}

Aff Engine::gn(const Tensor& x, const std::string& prefix, const Tensor* x2) {
    const int ct = x.c + (x2 ? x2->c : 0);
    KEEP_CHECK(ct % 32 == 0, "GroupNorm(32): %d channels", ct);
    const int cpg = ct / 32;
    Aff a;
    const float* g = warr(prefix + ".weight");
    const float* b = warr(prefix + ".bias");
    const int hw = x.h * x.w;
    if (x.aff && !x2) {   // the producing layer's split-K reduce already wrote this norm's affine
        KEEP_CHECK(x.aff_gamma == g, "groupnorm %s: the producer finalized its statistics for another norm", prefix.c_str());
        if (plan_) plan_->push_back("groupnorm n=" + std::to_string(x.n) + " hw=" + std::to_string(hw) + " c=" + std::to_string(ct) + " fused=2");
        Tensor& xm = const_cast<Tensor&>(x);
        a.scale = xm.aff;
        a.shift = a.scale + (size_t)x.n * ct;
        xm.aff = nullptr;
        if (xm.gn_part) { ar_->free(xm.gn_part); xm.gn_part = nullptr; }
        return a;
    }
    a.scale = (float*)ar_->alloc((size_t)2 * x.n * ct * sizeof(float));
    a.shift = a.scale + (size_t)x.n * ct;
    if (x.gn_part && !x2) {   // the producing kernel left the statistics: finalize only
        if (plan_) plan_->push_back("groupnorm n=" + std::to_string(x.n) + " hw=" + std::to_string(hw) + " c=" + std::to_string(ct) + " fused=1");
        if (!ar_->dry()) {
            gn_finalize_parts(x.gn_part, x.n, x.gn_P, hw, ct, 1e-6f, g, b, a.scale, a.shift, s_);
            launches_ += 1;
        }
        Tensor& xm = const_cast<Tensor&>(x);   // single consumer: release the slots (stream-ordered reuse is safe)
        ar_->free(xm.gn_part);
        xm.gn_part = nullptr;
        return a;
    }
    if (plan_) plan_->push_back("groupnorm n=" + std::to_string(x.n) + " hw=" + std::to_string(hw) + " c=" + std::to_string(ct));
    double* scratch = (double*)ar_->alloc(gn_scratch_doubles(x.n, hw, std::max(x.c, x2 ? x2->c : 0)) * sizeof(double));
    if (!ar_->dry()) {
        groupnorm_affine(x.p, x.dt, x.n, hw, x.c, cpg, 1e-6f, g, b, a.scale, a.shift, ct, 0, scratch, s_, gn_tickets_ ? gn_tickets_ + ((s_ == side_ && side_) ? gn_ticket_count() : 0) : nullptr);
        launches_ += 1;
        if (x2) {
            KEEP_CHECK(x.c % cpg == 0, "GroupNorm over concat: group straddles the sources");
            groupnorm_affine(x2->p, x2->dt, x2->n, hw, x2->c, cpg, 1e-6f, g, b, a.scale, a.shift, ct, x.c, scratch, s_, gn_tickets_ ? gn_tickets_ + ((s_ == side_ && side_) ? gn_ticket_count() : 0) : nullptr);
            launches_ += 1;
        }
    }
    ar_->free(scratch);
    return a;
}

Aff Engine::inorm(const Tensor& x) {
    Aff a;
    a.scale = (float*)ar_->alloc((size_t)2 * x.n * x.c * sizeof(float));
    a.shift = a.scale + (size_t)x.n * x.c;
    const int hw = x.h * x.w;
    double* scratch = (double*)ar_->alloc(gn_scratch_doubles(x.n, hw, x.c) * sizeof(double));
    if (!ar_->dry()) {
        groupnorm_affine(x.p, x.dt, x.n, hw, x.c, 1, 1e-5f, nullptr, nullptr, a.scale, a.shift, x.c, 0, scratch, s_, gn_tickets_ ? gn_tickets_ + ((s_ == side_ && side_) ? gn_ticket_count() : 0) : nullptr);
        launches_ += 2;
    }
    ar_->free(scratch);
    return a;
}

Tensor Engine::ln(const Tensor& x, const std::string& prefix, const Tensor* res, const float* add2, int add2_rows, Tensor* out2) {
    KEEP_CHECK(x.dt == F32, "layernorm expects fp32 tokens");
    if (plan_) plan_->push_back("layernorm rows=" + std::to_string(x.rows()) + " c=" + std::to_string(x.c));
    Tensor out = talloc(x.n, x.h, x.w, x.c, F32);
    if (out2) *out2 = talloc(x.n, x.h, x.w, x.c, F32);
    if (!ar_->dry()) {
        layernorm(x.f(), (int)x.rows(), x.c, warr(prefix + ".weight"), warr(prefix + ".bias"), 1e-5f, res ? res->f() : nullptr,
                  out.f(), add2, add2_rows, out2 ? out2->f() : nullptr, s_);
        launches_ += 1;
    }
    return out;
}

// vqgan_arch.py:170-181
Tensor Engine::res_block(const Tensor& x, const std::string& p, const Tensor* x2, bool out_stats, const std::string& out_norm) {
    Aff a1 = gn(x, p + ".norm1", x2);
    ConvOpt o1;
    o1.pad(1); o1.pre = &a1; o1.pre_act = ACT_SWISH; o1.in1 = x2; o1.want_stats = true; o1.stats_norm = p + ".norm2";   // h feeds norm2
    Tensor h = conv(x, p + ".conv1", o1);
    afree(a1);
    Aff a2 = gn(h, p + ".norm2");
    Tensor skip = x;
    const bool own_skip = has(p + ".conv_out.weight");
    if (own_skip) {
        ConvOpt os;
        os.in1 = x2;
        skip = conv(x, p + ".conv_out", os);
    } else {
        KEEP_CHECK(!x2, "res_block: concat input needs conv_out");
    }
    ConvOpt o2;
    o2.pad(1); o2.pre = &a2; o2.pre_act = ACT_SWISH; o2.res = &skip; o2.want_stats = out_stats; o2.stats_norm = out_norm;
    Tensor out = conv(h, p + ".conv2", o2);
    afree(a2);
    tfree(h);
    if (own_skip) tfree(skip);
    return out;
}

// C[z] = alpha * A[z] (M x K) * B[z]^T, B[z] = (N x K) strided view, on the tcgen05 kernel: B is packed into per-batch
// "weight" panel sets (tc_pack_matrix) and the batch rides on the kernel's image index.  fp32 in / out.
Tensor Engine::gemm_nt_tc(const float* A, int nb, int M, int K, const float* B, long long b_bstride, int ld_n, int ld_k, int N,
                          float alpha, int lda) {
    const long long m_tiles = (long long)nb * cdiv(M, 128);
    const int bn = tc_pick_bn(N, m_tiles, tc_passes_);
    if (plan_) plan_->push_back("gemm nb=" + std::to_string(nb) + " M=" + std::to_string(M) + " K=" + std::to_string(K) + " N=" +
                                std::to_string(N) + " kernel=tcgen05");
    const size_t per = tc_pack_matrix(nullptr, 0, 0, 0, nb, N, K, bn, tc_passes_, 1.0f, nullptr, s_);
    __half* panels = (__half*)ar_->alloc(per * nb * sizeof(__half));
    Tensor out = talloc(nb, M, 1, N, F32);
    if (!ar_->dry()) {
        tc_pack_matrix(B, b_bstride, ld_n, ld_k, nb, N, K, bn, tc_passes_, alpha, panels, s_);
        ConvArgs a;
        a.in0 = A; a.in0_dt = F32; a.c0 = K; a.ld0 = lda;   // (lda > 0: A[z] = M rows of a wider matrix, batch stride M * lda)
        a.n = nb; a.h = M; a.w = 1; a.up = 1;
        a.kh = 1; a.kw = 1; a.stride = 1; a.cout = N; a.ho = M; a.wo = 1;
        a.out = out.p; a.out_dt = F32;
        a.wt_img_stride = (long long)per;
        Prof pr;
        if (profile_) {
            pr.a = get_event(); pr.b = get_event();
            pr.flops = 2.0 * nb * (double)M * N * K;
            pr.bytes = 4.0 * nb * ((double)M * K + (double)N * K + (double)M * N);
            pr.tag = 1; pr.m = nb * M; pr.k = K; pr.n = N; pr.kh = 11; pr.splitk = 1; pr.bn = bn;
            CUDA_CHECK(cudaEventRecord(pr.a, s_));
        }
        const int grid_cap = (s_ == side_ && side_) ? side_sms_ : main_cap_;
        conv2d_tc(a, panels, bn, tc_passes_, 1, nullptr, grid_cap, s_);
        launches_ += 2;
        if (profile_) { CUDA_CHECK(cudaEventRecord(pr.b, s_)); prof_.push_back(pr); }
    }
    ar_->free(panels);
    return out;
}

// generic multi-head attention: scores -> softmax -> PV, all fp32
Tensor Engine::mha(const float* q, int ldq, long long sq, const float* k, int ldk, long long sk, const float* v, int ldv,
                   long long sv, int nb, int Lq, int Lk, int heads, int dh, float scale) {
    if (plan_) plan_->push_back("attention nb=" + std::to_string(nb) + " Lq=" + std::to_string(Lq) + " Lk=" + std::to_string(Lk) +
                                " heads=" + std::to_string(heads) + " dh=" + std::to_string(dh));
    Tensor S = talloc(nb * heads, Lq, 1, Lk, F32);
    Tensor O = talloc(nb, Lq, 1, heads * dh, F32);
    if (!ar_->dry()) {
        BGemmArgs g;
        g.A = q; g.B = k; g.C = S.f();
        g.M = Lq; g.N = Lk; g.K = dh; g.lda = ldq; g.ldb = ldk; g.ldc = Lk; g.transB = 1; g.alpha = scale;
        g.nz0 = nb; g.nz1 = heads;
        g.sA[0] = sq; g.sA[1] = dh; g.sB[0] = sk; g.sB[1] = dh;
        g.sC[0] = (long long)heads * Lq * Lk; g.sC[1] = (long long)Lq * Lk;
        bgemm_simt(g, s_);
        softmax_rows(S.f(), (long long)nb * heads * Lq, Lk, nullptr, 1, Lq, s_);
        BGemmArgs h;
        h.A = S.f(); h.B = v; h.C = O.f();
        h.M = Lq; h.N = dh; h.K = Lk; h.lda = Lk; h.ldb = ldv; h.ldc = heads * dh; h.transB = 0;
        h.nz0 = nb; h.nz1 = heads;
        h.sA[0] = (long long)heads * Lq * Lk; h.sA[1] = (long long)Lq * Lk;
        h.sB[0] = sv; h.sB[1] = dh;
        h.sC[0] = (long long)Lq * heads * dh; h.sC[1] = dh;
        bgemm_simt(h, s_);
        launches_ += 3;
    }
    tfree(S);
    return O;
}

// Multi-head attention with every contraction on the tensor cores (single batch; q / k / v token-major with the heads side by
// side).  Scores: ONE GEMM over K = heads * dh channels run in split-K = heads mode WITHOUT the reduce -- the K split of head h
// covers exactly that head's channel blocks, so partial tile h is Q_h K_h^T (the key matrix is packed as the B operand with
// the softmax scale folded in).  P V: the heads ride on the kernel's image index with per-image B panels (V_h^T as a strided
// view), then one permute back to token-major.  Used where the SIMT GEMMs were the cost: CFA at 32^2 (1024 x 1024 x 4 x 256,
// keep_arch.py:200-241,519-541: 72 + 68 us per frame on the CUDA cores).
Tensor Engine::mha_tc(const float* q, const float* k, const float* v, int Lq, int Lk, int heads, int dh, float scale) {
    const int inner = heads * dh;
    KEEP_CHECK(dh % tc_cb(tc_passes_) == 0 && Lk % 16 == 0 && Lq % 8 == 0, "mha_tc: unsupported shape (Lq %d, Lk %d, dh %d)", Lq, Lk, dh);
    if (plan_) plan_->push_back("attention nb=1 Lq=" + std::to_string(Lq) + " Lk=" + std::to_string(Lk) + " heads=" + std::to_string(heads) +
                                " dh=" + std::to_string(dh) + " kernel=tcgen05");
    const int grid_cap = (s_ == side_ && side_) ? side_sms_ : main_cap_;
    const int bn = tc_pick_bn(Lk, cdiv(Lq, 128), tc_passes_);
    const size_t per = tc_pack_matrix(nullptr, 0, 0, 0, 1, Lk, inner, bn, tc_passes_, 1.0f, nullptr, s_);
    __half* panels = (__half*)ar_->alloc(per * sizeof(__half));
    Tensor S = talloc(heads, Lq, 1, Lk, F32);
    if (!ar_->dry()) {
        tc_pack_matrix(k, 0, inner, 1, 1, Lk, inner, bn, tc_passes_, scale, panels, s_);
        ConvArgs a;
        a.in0 = q; a.in0_dt = F32; a.c0 = inner;
        a.n = 1; a.h = Lq; a.w = 1; a.up = 1;
        a.kh = 1; a.kw = 1; a.stride = 1; a.cout = Lk; a.ho = Lq; a.wo = 1;
        a.out = S.p; a.out_dt = F32;
        a.wt_img_stride = (long long)per;
        a.splitk = heads; a.partial = S.f(); a.no_reduce = 1;
        conv2d_tc(a, panels, bn, tc_passes_, heads, S.f(), grid_cap, s_);
        softmax_rows(S.f(), (long long)heads * Lq, Lk, nullptr, 1, Lq, s_);
        launches_ += 3;
    }
    ar_->free(panels);
    std::vector<std::string>* plan_saved = plan_;
    plan_ = nullptr;                                                            // (the "attention" plan line above covers both contractions)
    Tensor Oh = gemm_nt_tc(S.f(), heads, Lq, Lk, v, dh, 1, inner, dh, 1.0f);   // (heads, Lq, dh); B[h] = V_h^T, a strided view
    plan_ = plan_saved;
    tfree(S);
    Tensor O = talloc(1, Lq, 1, inner, F32);
    if (!ar_->dry()) { heads_to_tokens(Oh.f(), O.f(), heads, Lq, dh, s_); launches_ += 1; }
    tfree(Oh);
    return O;
}

// vqgan_arch.py:219-243
Tensor Engine::attn_block(const Tensor& x, const std::string& p, bool out_stats, const std::string& out_norm) {
    Aff a = gn(x, p + ".norm");
    ConvOpt o;
    o.pre = &a; o.out_dt = F32;
    Tensor qkv = conv(x, p + ".qkv", o);
    afree(a);
    const int L = x.h * x.w, C = x.c;
    // both contractions on the tcgen05 GEMM kernel (q / k / v are column slices of the fused qkv matrix: strided A operand, strided
    // B packs), scores materialised (n x L x L fp32, 256 KB per image at 16^2).  KEEP_ATTNBLOCK_TC=0: the CUDA-core batched GEMMs of
    // round 1 (measured equal in time: 177.3 vs 177.0 frames/s, profiles/r2_experiments.md)
    static const bool attn_tc = !(getenv("KEEP_ATTNBLOCK_TC") && getenv("KEEP_ATTNBLOCK_TC")[0] == '0');
    Tensor O;
    if (attn_tc && (flags_ & KEEP_FLAG_TCGEN05) && tc_passes_ == 3 && L % 128 == 0 && C % 32 == 0) {
        if (plan_) plan_->push_back("attention nb=" + std::to_string(x.n) + " Lq=" + std::to_string(L) + " Lk=" + std::to_string(L) +
                                    " heads=1 dh=" + std::to_string(C) + " kernel=tcgen05");
        std::vector<std::string>* plan_saved = plan_;
        plan_ = nullptr;                                    // (the "attention" line covers both contractions)
        const long long bs = (long long)L * 3 * C;
        Tensor S = gemm_nt_tc(qkv.f(), x.n, L, C, qkv.f() + C, bs, 3 * C, 1, L, 1.0f / sqrtf((float)C), 3 * C);
        if (!ar_->dry()) { softmax_rows(S.f(), (long long)x.n * L, L, nullptr, 1, L, s_); launches_ += 1; }
        O = gemm_nt_tc(S.f(), x.n, L, L, qkv.f() + 2 * C, bs, 1, 3 * C, C, 1.0f);
        plan_ = plan_saved;
        tfree(S);
    } else {
        O = mha(qkv.f(), 3 * C, (long long)L * 3 * C, qkv.f() + C, 3 * C, (long long)L * 3 * C, qkv.f() + 2 * C, 3 * C,
                (long long)L * 3 * C, x.n, L, L, 1, C, 1.0f / sqrtf((float)C));
    }
    tfree(qkv);
    O.h = x.h; O.w = x.w;
    ConvOpt po;
    po.res = &x; po.want_stats = out_stats; po.stats_norm = out_norm;
    Tensor out = conv(O, p + ".proj_out", po);
    tfree(O);
    return out;
}

// Encoder programme (vqgan_arch.py:246-292) for nf 64, ch_mult [1,2,2,4,4,8], 2 res blocks, attention at 16
static const char* kEncProg[25] = {"conv", "res", "res", "down", "res", "res", "down", "res", "res", "down", "res", "res", "down",
                                   "res", "res", "down", "res", "attn", "res", "attn", "res", "attn", "res", "norm", "conv"};
static const char* kGenProg[25] = {"conv", "res", "attn", "res", "res", "attn", "res", "attn", "up", "res", "res", "up", "res",
                                   "res", "up", "res", "res", "up", "res", "res", "up", "res", "res", "norm", "conv"};

Tensor Engine::encoder(const Tensor& img, const std::string& p, const std::function<void(int, const Tensor&)>& tap) {
    Tensor x = img;
    bool own = false;
    Aff pend;  // pending GroupNorm (norm_out) to fuse into the last conv
    for (int i = 0; i < 25; ++i) {
        const std::string bp = p + ".blocks." + std::to_string(i);
        const std::string kind = kEncProg[i];
        // the next block starts with a GroupNorm of this block's output (res: norm1, attn: norm, norm_out)
        const std::string next = i + 1 < 25 ? kEncProg[i + 1] : "";
        const bool next_gn = next == "res" || next == "attn" || next == "norm";
        const std::string nbp = p + ".blocks." + std::to_string(i + 1);
        const std::string next_norm = !next_gn ? "" : (next == "res" ? nbp + ".norm1" : (next == "attn" ? nbp + ".norm" : nbp));
        Tensor y;
        if (kind == "conv") {
            ConvOpt o;
            o.pad(1); o.want_stats = next_gn; o.stats_norm = next_norm;
            if (i == 24) { o.pre = &pend; o.out_dt = F32; }
            y = conv(x, bp, o);
            if (i == 24) afree(pend);
        } else if (kind == "res") {
            y = res_block(x, bp, nullptr, next_gn, next_norm);
        } else if (kind == "attn") {
            y = attn_block(x, bp, next_gn, next_norm);
        } else if (kind == "down") {  // vqgan_arch.py:135-139
            ConvOpt o;
            o.stride = 2; o.pad_b = 1; o.pad_r = 1; o.want_stats = next_gn; o.stats_norm = next_norm;
            y = conv(x, bp + ".conv", o);
        } else {  // norm_out: statistics only; apply is fused into the next conv
            pend = gn(x, bp);
            if (tap) tap(i, x);
            continue;
        }
        if (own) tfree(x);
        x = y;
        own = true;
        if (tap) tap(i, x);
    }
    return x;
}

// =============================================================================================
// GMFlow (gmflow/*.py; 1 scale, swin splits 2, global correlation + global propagation)
// =============================================================================================
// gmflow/backbone.py:8-36.  x_aff != null: x is a raw conv output with a pending InstanceNorm+ReLU.
void Engine::gm_resblock(Tensor& x, Aff* x_aff, const std::string& p, int stride) {
    ConvOpt o1;
    o1.pad(1); o1.stride = stride; o1.pre = x_aff; o1.pre_act = x_aff ? ACT_RELU : ACT_NONE; o1.out_dt = F32;
    Tensor y1 = conv(x, p + ".conv1", o1);
    Aff a1 = inorm(y1);
    ConvOpt o2;
    o2.pad(1); o2.pre = &a1; o2.pre_act = ACT_RELU; o2.out_dt = F32;
    Tensor y2 = conv(y1, p + ".conv2", o2);
    afree(a1);
    tfree(y1);
    Aff a2 = inorm(y2);
    Tensor out = talloc(y2.n, y2.h, y2.w, y2.c, F32);
    EwArgs e;
    Tensor d;
    Aff a3;
    const bool ds = has(p + ".downsample.0.weight");
    if (ds) {
        ConvOpt od;
        od.stride = stride; od.pre = x_aff; od.pre_act = x_aff ? ACT_RELU : ACT_NONE; od.out_dt = F32;
        d = conv(x, p + ".downsample.0", od);
        a3 = inorm(d);
        e.A = d.p; e.a_dt = d.dt; e.sa = a3.scale; e.ba = a3.shift; e.act_a = ACT_NONE;
    } else {
        e.A = x.p; e.a_dt = x.dt;
        if (x_aff) { e.sa = x_aff->scale; e.ba = x_aff->shift; e.act_a = ACT_RELU; }
    }
    e.B = y2.p; e.b_dt = y2.dt; e.sb = a2.scale; e.bb = a2.shift; e.act_b = ACT_RELU;
    e.act_o = ACT_RELU;
    e.out = out.p; e.o_dt = out.dt; e.n = out.n; e.hw = out.h * out.w; e.c = out.c;
    if (!ar_->dry()) { elementwise(e, s_); launches_ += 1; }
    if (ds) { afree(a3); tfree(d); }
    afree(a2);
    tfree(y2);
    tfree(x);
    x = out;
}

// gmflow/transformer.py:147-185 on tokens (nimg, 4096, 128); windows 2x2, optional half-window shift
void Engine::gm_layer(Tensor& src, const Tensor& tgt, const std::string& p, int nimg, bool shift, bool ffn) {
    const int C = 128, H = 64, Wd = 64, k = 2, L = 1024;
    // KEEP_GM_FUSE_QKV=1 (experiment, default off until measured): q|k|v (self) / k|v (cross) as one GEMM over fused weights
    static const bool fuse = getenv("KEEP_GM_FUSE_QKV") != nullptr && atoi(getenv("KEEP_GM_FUSE_QKV")) != 0;
    const bool fuse_here = fuse && has(p + ".qkv_proj.weight") && has(p + ".kv_proj.weight");
    Tensor q, kk, v;           // q / k / v projections: separate tensors, or strided views into one fused GEMM output
    const float *qp, *kp, *vp;
    int ldq = C, ldkv = C;
    if (fuse_here && src.p == tgt.p) {
        q = linear(src, p + ".qkv_proj");
        qp = q.f(); kp = q.f() + C; vp = q.f() + 2 * C; ldq = ldkv = 3 * C;
    } else if (fuse_here) {
        q = linear(src, p + ".q_proj");
        kk = linear(tgt, p + ".kv_proj");
        qp = q.f(); kp = kk.f(); vp = kk.f() + C; ldkv = 2 * C;
    } else {
        q = linear(src, p + ".q_proj"); kk = linear(tgt, p + ".k_proj"); v = linear(tgt, p + ".v_proj");
        qp = q.f(); kp = kk.f(); vp = v.f();
    }
    const int sh = shift ? 16 : 0;
    static const bool fused_attn = !(getenv("KEEP_FUSED_ATTN") && getenv("KEEP_FUSED_ATTN")[0] == '0');
    Tensor O;
    if ((flags_ & KEEP_FLAG_TCGEN05) && tc_passes_ == 3 && fused_attn && attention_tc_eligible(L, L, C)) {
        // fused window attention (attn_tcgen05.cu): scores and probabilities stay in TMEM / shared memory; the window
        // partition with its cyclic shift and the merge are index math inside the kernel (gmflow/transformer.py:78-103)
        if (plan_) plan_->push_back("attention nb=" + std::to_string(nimg * 4) + " Lq=" + std::to_string(L) + " Lk=" + std::to_string(L) +
                                    " heads=1 dh=" + std::to_string(C) + " kernel=tcgen05_fused");
        O = talloc(nimg, H * Wd, 1, C, F32);
        // a window's 8 query tiles share its K / V: convert them once (pack kernel) and stream the stages by TMA (KEEP_ATTN_PACK=0:
        // every query tile converts them itself)
        static const bool attn_pack = !(getenv("KEEP_ATTN_PACK") && getenv("KEEP_ATTN_PACK")[0] == '0');
        void* pack = attn_pack ? ar_->alloc(attention_tc_pack_bytes(nimg * 4, L, C)) : nullptr;
        if (!ar_->dry()) {
            attention_tc(qp, ldq, (long long)H * Wd * ldq, kp, ldkv, (long long)H * Wd * ldkv, vp, ldkv, (long long)H * Wd * ldkv, O.f(), C,
                         (long long)H * Wd * C, nimg * 4, L, L, C, 1.0f / sqrtf((float)C), shift ? region8_ : nullptr, 4, s_, k, 32, Wd, sh, 1, pack);
            launches_ += pack ? 2 : 1;
        }
        if (pack) ar_->free(pack);
        tfree(q);
        if (kk.p) tfree(kk);
        if (v.p) tfree(v);
    } else {
    Tensor qw = talloc(nimg * 4, L, 1, C, F32), kw = talloc(nimg * 4, L, 1, C, F32), vw = talloc(nimg * 4, L, 1, C, F32);
    if (!ar_->dry()) {
        window_partition(qp, qw.f(), nimg, H, Wd, C, k, sh, sh, ldq, s_);
        window_partition(kp, kw.f(), nimg, H, Wd, C, k, sh, sh, ldkv, s_);
        window_partition(vp, vw.f(), nimg, H, Wd, C, k, sh, sh, ldkv, s_);
        launches_ += 3;
    }
    tfree(q);
    if (kk.p) tfree(kk);
    if (v.p) tfree(v);
    Tensor S, Ow;
    if (flags_ & KEEP_FLAG_TCGEN05) {
        // window attention on the tensor cores: S = (Q K^T)/sqrt(C), softmax (+ shift mask), O = P V
        S = gemm_nt_tc(qw.f(), nimg * 4, L, C, kw.f(), (long long)L * C, C, 1, L, 1.0f / sqrtf((float)C));
        if (!ar_->dry()) {
            softmax_rows(S.f(), (long long)nimg * 4 * L, L, shift ? region_ : nullptr, 4, L, s_);
            launches_ += 1;
        }
        Ow = gemm_nt_tc(S.f(), nimg * 4, L, L, vw.f(), (long long)L * C, 1, C, C, 1.0f);   // B = V^T as a strided view
    } else {
        S = talloc(nimg * 4, L, 1, L, F32);
        Ow = talloc(nimg * 4, L, 1, C, F32);
        if (!ar_->dry()) {
            BGemmArgs g;
            g.A = qw.f(); g.B = kw.f(); g.C = S.f();
            g.M = L; g.N = L; g.K = C; g.lda = C; g.ldb = C; g.ldc = L; g.transB = 1; g.alpha = 1.0f / sqrtf((float)C);
            g.nz0 = nimg * 4;
            g.sA[0] = (long long)L * C; g.sB[0] = (long long)L * C; g.sC[0] = (long long)L * L;
            bgemm_simt(g, s_);
            softmax_rows(S.f(), (long long)nimg * 4 * L, L, shift ? region_ : nullptr, 4, L, s_);
            BGemmArgs h;
            h.A = S.f(); h.B = vw.f(); h.C = Ow.f();
            h.M = L; h.N = C; h.K = L; h.lda = L; h.ldb = C; h.ldc = C; h.transB = 0;
            h.nz0 = nimg * 4;
            h.sA[0] = (long long)L * L; h.sB[0] = (long long)L * C; h.sC[0] = (long long)L * C;
            bgemm_simt(h, s_);
            launches_ += 3;
        }
    }
    tfree(S); tfree(qw); tfree(kw); tfree(vw);
    O = talloc(nimg, H * Wd, 1, C, F32);
    if (!ar_->dry()) { window_merge(Ow.f(), O.f(), nimg, H, Wd, C, k, sh, sh, s_); launches_ += 1; }
    tfree(Ow);
    }
    Tensor m = linear(O, p + ".merge");
    tfree(O);
    Tensor out;
    if (!ffn) {
        out = ln(m, p + ".norm1", &src);
        tfree(m);
    } else {
        Tensor m1 = ln(m, p + ".norm1");
        tfree(m);
        ConvOpt o;
        o.in1 = &m1; o.act = ACT_GELU; o.out_dt = F32;   // mlp.0 on cat([source, message]) without materialising the concat
        Tensor h = conv(src, convw(p + ".mlp.0"), o);
        tfree(m1);
        Tensor m2 = linear(h, p + ".mlp.2");
        tfree(h);
        out = ln(m2, p + ".norm2", &src);
        tfree(m2);
    }
    tfree(src);
    src = out;
}

// x_nchw: (T,3,512,512) fp32 in [-1,1]; flows: (T-1,512,512,2), pair i = flow(frame i+1 -> frame i)  (keep_arch.py:976-986)
void Engine::gmflow(const float* x_nchw, int T, float* flows, int p_lo, int p_hi) {
    const std::string P = "flownet.model";
    const int HW = 512 * 512;
    const int chunk = flow_chunk();
    const int p_end = p_hi < 0 ? T - 1 : std::min(T - 1, p_hi);   // pairs [p_lo, p_end), p_lo a multiple of the chunk size
    for (int p0 = p_lo; p0 < p_end; p0 += chunk) {
        const int np = std::min(chunk, T - 1 - p0);
        // images: [img0 = frames p0+1 .. p0+np | img1 = frames p0 .. p0+np-1], ImageNet-normalised NHWC
        Tensor img = talloc(2 * np, 512, 512, 3, F32);
        if (!ar_->dry()) {
            nchw_to_nhwc(x_nchw + (size_t)(p0 + 1) * 3 * HW, img.p, F32, np, 3, 512, 512, 1, s_);
            nchw_to_nhwc(x_nchw + (size_t)p0 * 3 * HW, (float*)img.p + (size_t)np * HW * 3, F32, np, 3, 512, 512, 1, s_);
            launches_ += 2;
        }
        // backbone (gmflow/backbone.py:101-117)
        ConvOpt o;
        o.stride = 2; o.pad(3); o.out_dt = F32;
        Tensor x = conv(img, P + ".backbone.conv1", o);
        tfree(img);
        Aff a = inorm(x);
        gm_resblock(x, &a, P + ".backbone.layer1.0", 1);
        afree(a);
        gm_resblock(x, nullptr, P + ".backbone.layer1.1", 1);
        gm_resblock(x, nullptr, P + ".backbone.layer2.0", 2);
        gm_resblock(x, nullptr, P + ".backbone.layer2.1", 1);
        gm_resblock(x, nullptr, P + ".backbone.layer3.0", 2);
        gm_resblock(x, nullptr, P + ".backbone.layer3.1", 1);
        ConvOpt oc;
        oc.out_dt = F32;
        Tensor feat = conv(x, P + ".backbone.conv2", oc);   // (2np, 64, 64, 128)
        tfree(x);
        if (!ar_->dry()) { add_window_sine_pos(feat.f(), 2 * np, 64, 64, 128, 2, s_); launches_ += 1; }
        // transformer (gmflow/transformer.py:273-322): c0 = [f0; f1], c1 = [f1; f0]
        Tensor c0 = feat;
        c0.h = 4096; c0.w = 1;
        Tensor c1 = talloc(2 * np, 4096, 1, 128, F32);
        const size_t half = (size_t)np * 4096 * 128;
        auto swap_into_c1 = [&]() {
            if (ar_->dry()) return;
            CUDA_CHECK(cudaMemcpyAsync(c1.f(), c0.f() + half, half * sizeof(float), cudaMemcpyDeviceToDevice, s_));
            CUDA_CHECK(cudaMemcpyAsync(c1.f() + half, c0.f(), half * sizeof(float), cudaMemcpyDeviceToDevice, s_));
        };
        swap_into_c1();
        for (int l = 0; l < 6; ++l) {
            const std::string lp = P + ".transformer.layers." + std::to_string(l);
            const bool shift = (l % 2) == 1;
            gm_layer(c0, c0, lp + ".self_attn", 2 * np, shift, false);
            gm_layer(c0, c1, lp + ".cross_attn_ffn", 2 * np, shift, true);
            swap_into_c1();
        }
        tfree(c1);
        Tensor f0 = c0;  // view: first np images
        f0.n = np;
        const float* f1 = c0.f() + half;
        // global correlation softmax (gmflow/matching.py:7-36)
        const bool tcg = (flags_ & KEEP_FLAG_TCGEN05) != 0;
        Tensor S;
        Tensor flow = talloc(np, 64, 64, 2, F32);
        if (tcg) {
            S = gemm_nt_tc(f0.f(), np, 4096, 128, f1, 4096LL * 128, 128, 1, 4096, 1.0f / sqrtf(128.0f));
        } else {
            S = talloc(np, 4096, 1, 4096, F32);
            if (!ar_->dry()) {
                BGemmArgs g;
                g.A = f0.f(); g.B = f1; g.C = S.f();
                g.M = 4096; g.N = 4096; g.K = 128; g.lda = 128; g.ldb = 128; g.ldc = 4096; g.transB = 1;
                g.alpha = 1.0f / sqrtf(128.0f);
                g.nz0 = np;
                g.sA[0] = 4096LL * 128; g.sB[0] = 4096LL * 128; g.sC[0] = 4096LL * 4096;
                bgemm_simt(g, s_);
                launches_ += 1;
            }
        }
        if (!ar_->dry()) {
            softmax_expect2(S.f(), (long long)np * 4096, 4096, 4096, grid64_, 0, grid64_, flow.f(), s_);
            launches_ += 1;
        }
        // flow propagation (gmflow/transformer.py:343-374): q = q_proj(f0), k = k_proj(q)
        Tensor q = linear(f0, P + ".feature_flow_attn.q_proj");
        Tensor kq = linear(q, P + ".feature_flow_attn.k_proj");
        Tensor flow2 = talloc(np, 64, 64, 2, F32);
        if (tcg) {
            tfree(S);
            S = gemm_nt_tc(q.f(), np, 4096, 128, kq.f(), 4096LL * 128, 128, 1, 4096, 1.0f / sqrtf(128.0f));
        } else if (!ar_->dry()) {
            BGemmArgs g;
            g.A = q.f(); g.B = kq.f(); g.C = S.f();
            g.M = 4096; g.N = 4096; g.K = 128; g.lda = 128; g.ldb = 128; g.ldc = 4096; g.transB = 1;
            g.alpha = 1.0f / sqrtf(128.0f);
            g.nz0 = np;
            g.sA[0] = 4096LL * 128; g.sB[0] = 4096LL * 128; g.sC[0] = 4096LL * 4096;
            bgemm_simt(g, s_);
            launches_ += 1;
        }
        if (!ar_->dry()) {
            softmax_expect2(S.f(), (long long)np * 4096, 4096, 4096, flow.f(), 4096LL * 2, nullptr, flow2.f(), s_);
            launches_ += 1;
        }
        tfree(S); tfree(q); tfree(kq); tfree(flow);
        // convex x8 upsampling (gmflow/gmflow.py:74-88): conv3x3 on cat(flow, feature0) -> ReLU -> 1x1 -> 576 logits
        Tensor f0m = f0;
        f0m.h = 64; f0m.w = 64;
        ConvOpt ou;
        ou.pad(1); ou.act = ACT_RELU; ou.out_dt = F32;
        Tensor u;
        if (tcg && has(P + ".upsampler.0.perm136.weight")) {   // tensor-core path: [feature | flow padded to 8 channels]
            Tensor flow8 = talloc(np, 64, 64, 8, F32);
            if (!ar_->dry()) { concat2(flow2.f(), 2, nullptr, 6, flow8.f(), (long long)np * 4096, s_); launches_ += 1; }
            ConvW cw = convw(P + ".upsampler.0.perm136");
            cw.b = convw(P + ".upsampler.0").b;
            ou.in1 = &flow8;
            u = conv(f0m, cw, ou);
            tfree(flow8);
        } else {
            ou.in1 = &f0m;
            u = conv(flow2, convw(P + ".upsampler.0"), ou);
        }
        ConvOpt om;
        om.out_dt = F32;
        Tensor mask = conv(u, P + ".upsampler.2", om);
        tfree(u);
        if (!ar_->dry()) {
            convex_upsample8(mask.f(), flow2.f(), flows + (size_t)p0 * HW * 2, np, 64, 64, s_);
            launches_ += 1;
        }
        tfree(mask); tfree(flow2); tfree(c0);
        if (!ar_->dry() && s_ == side_ && side_) CUDA_CHECK(cudaEventRecord(ev_flow_[ev_flow_base_ + p0 / chunk], side_));
    }
}

// =============================================================================================
// Kalman gain estimator (keep_arch.py:801-821, BasicTransformerBlock :640-682)
// =============================================================================================
Tensor Engine::kalman_gains(const Tensor& z_codes, int T) {
    const int L = 256, C = 256, heads = 8, dh = 48, inner = heads * dh;
    const float scale = 1.0f / sqrtf((float)dh);
    Tensor h = talloc(1, T * L, 1, C, F32);
    if (!ar_->dry()) CUDA_CHECK(cudaMemcpyAsync(h.p, z_codes.p, h.bytes(), cudaMemcpyDeviceToDevice, s_));
    for (int blk = 0; blk < 3; ++blk) {
        const std::string p = "kalman_filter.uncertainty_estimator." + std::to_string(blk);
        // sparse-causal attention: keys/values = [frame 0 || frame i-1]
        Tensor hn = ln(h, p + ".norm1");
        Tensor q = linear(hn, p + ".attn1.to_q"), k = linear(hn, p + ".attn1.to_k"), v = linear(hn, p + ".attn1.to_v");
        tfree(hn);
        Tensor k2 = talloc(1, T * 2 * L, 1, inner, F32), v2 = talloc(1, T * 2 * L, 1, inner, F32);
        if (!ar_->dry()) {
            sparse_causal_gather(k.f(), k2.f(), 1, T, L, inner, s_);
            sparse_causal_gather(v.f(), v2.f(), 1, T, L, inner, s_);
            launches_ += 2;
        }
        tfree(k); tfree(v);
        Tensor o = mha(q.f(), inner, (long long)L * inner, k2.f(), inner, 2LL * L * inner, v2.f(), inner, 2LL * L * inner, T, L,
                       2 * L, heads, dh, scale);
        tfree(q); tfree(k2); tfree(v2);
        Tensor h1 = linear(o, p + ".attn1.to_out.0", ACT_NONE, &h);
        tfree(o); tfree(h);
        // GEGLU feed-forward
        Tensor n3 = ln(h1, p + ".norm3");
        Tensor pr = linear(n3, p + ".ff.net.0.proj");
        tfree(n3);
        Tensor gg = talloc(1, T * L, 1, 4 * C, F32);
        if (!ar_->dry()) { geglu(pr.f(), gg.f(), T * L, 4 * C, s_); launches_ += 1; }
        tfree(pr);
        Tensor h2 = linear(gg, p + ".ff.net.2", ACT_NONE, &h1);
        tfree(gg); tfree(h1);
        // temporal attention over frames for every spatial token: batch = token, sequence = frame
        Tensor nt = ln(h2, p + ".norm_temp");
        Tensor qt = linear(nt, p + ".attn_temp.to_q"), kt = linear(nt, p + ".attn_temp.to_k"), vt = linear(nt, p + ".attn_temp.to_v");
        tfree(nt);
        Tensor St = talloc(L * heads, T, 1, T, F32);
        Tensor ot = talloc(1, T * L, 1, inner, F32);
        if (!ar_->dry()) {
            BGemmArgs g;
            g.A = qt.f(); g.B = kt.f(); g.C = St.f();
            g.M = T; g.N = T; g.K = dh; g.lda = L * inner; g.ldb = L * inner; g.ldc = T; g.transB = 1; g.alpha = scale;
            g.nz0 = L; g.nz1 = heads;
            g.sA[0] = inner; g.sA[1] = dh; g.sB[0] = inner; g.sB[1] = dh;
            g.sC[0] = (long long)heads * T * T; g.sC[1] = (long long)T * T;
            bgemm_simt(g, s_);
            softmax_rows(St.f(), (long long)L * heads * T, T, nullptr, 1, T, s_);
            BGemmArgs m;
            m.A = St.f(); m.B = vt.f(); m.C = ot.f();
            m.M = T; m.N = dh; m.K = T; m.lda = T; m.ldb = L * inner; m.ldc = L * inner; m.transB = 0;
            m.nz0 = L; m.nz1 = heads;
            m.sA[0] = (long long)heads * T * T; m.sA[1] = (long long)T * T;
            m.sB[0] = inner; m.sB[1] = dh; m.sC[0] = inner; m.sC[1] = dh;
            bgemm_simt(m, s_);
            launches_ += 3;
        }
        tfree(St); tfree(qt); tfree(kt); tfree(vt);
        h = linear(ot, p + ".attn_temp.to_out.0", ACT_NONE, &h2);
        tfree(ot); tfree(h2);
    }
    // kalman_gain_calculator: 3 ResBlocks, conv1x1 -> 1, sigmoid (keep_arch.py:766-772)
    Tensor x = h;
    x.n = T; x.h = 16; x.w = 16;
    const int save = adt_;
    adt_ = F32;
    for (int i = 0; i < 3; ++i) {
        Tensor y = res_block(x, "kalman_filter.kalman_gain_calculator." + std::to_string(i), nullptr, i < 2,
                             i < 2 ? "kalman_filter.kalman_gain_calculator." + std::to_string(i + 1) + ".norm1" : "");
        tfree(x);
        x = y;
    }
    ConvOpt o;
    o.act = ACT_SIGMOID; o.out_dt = F32;
    Tensor g = conv(x, "kalman_filter.kalman_gain_calculator.3", o);   // (T,16,16,1)
    adt_ = save;
    tfree(x);
    return g;
}

// =============================================================================================
// code-prediction transformer (keep_arch.py:1073-1089) -> quantised latent (1,16,16,256)
// =============================================================================================
Tensor Engine::code_transformer(const Tensor& z_hat, int frame) {
    const int L = 256, E = 512, heads = 8, dh = 64;
    const int nbt = z_hat.n;   // clips in lockstep (1 on the per-clip path): tokens of clip c are rows [c*L, (c+1)*L)
    Tensor zt = z_hat;
    zt.n = 1; zt.h = nbt * L; zt.w = 1;
    const float* pos = warr("position_emb");
    // every LayerNorm of the stack follows a split-K linear: its reduce kernel writes the normalised rows too (LnFuse)
    LnFuse pre;
    pre.prefix = "ft_layers.0.norm1"; pre.add2 = pos; pre.add2_rows = L;
    Tensor t = linear(zt, "feat_emb", ACT_NONE, nullptr, &pre);
    for (int l = 0; l < 9; ++l) {
        const std::string p = "ft_layers." + std::to_string(l);
        Tensor qk_in, tn;
        if (pre.done) { tn = pre.out; qk_in = pre.out2; }
        else tn = ln(t, p + ".norm1", nullptr, pos, L, &qk_in);
        Tensor qk = linear(qk_in, p + ".self_attn.in_proj_qk");
        Tensor v = linear(tn, p + ".self_attn.in_proj_v");
        tfree(qk_in); tfree(tn);
        // nn.MultiheadAttention 8 x 64 over 256 tokens (keep_arch.py:431-432): the fused tcgen05 attention kernel, heads as batch
        static const bool fused_mha = !(getenv("KEEP_FUSED_MHA") && getenv("KEEP_FUSED_MHA")[0] == '0');
        Tensor o;
        if ((flags_ & KEEP_FLAG_TCGEN05) && tc_passes_ == 3 && fused_mha && attention_tc_eligible(L, L, dh)) {
            if (plan_) plan_->push_back("attention nb=" + std::to_string(nbt) + " Lq=" + std::to_string(L) + " Lk=" + std::to_string(L) +
                                        " heads=" + std::to_string(heads) + " dh=" + std::to_string(dh) + " kernel=tcgen05_fused");
            o = talloc(nbt, L, 1, E, F32);
            if (!ar_->dry()) {
                attention_tc(qk.f(), 2 * E, (long long)L * 2 * E, qk.f() + E, 2 * E, (long long)L * 2 * E, v.f(), E, (long long)L * E, o.f(), E,
                             (long long)L * E, nbt * heads, L, L, dh, 1.0f / sqrtf((float)dh), nullptr, 1, s_, 0, 0, 0, 0, heads);
                launches_ += 1;
            }
        } else {
            o = mha(qk.f(), 2 * E, (long long)L * 2 * E, qk.f() + E, 2 * E, (long long)L * 2 * E, v.f(), E, (long long)L * E, nbt, L, L,
                    heads, dh, 1.0f / sqrtf((float)dh));
        }
        tfree(qk); tfree(v);
        LnFuse f2;
        f2.prefix = p + ".norm2";
        Tensor t1 = linear(o, p + ".self_attn.out_proj", ACT_NONE, &t, &f2);
        tfree(o); tfree(t);
        Tensor n2 = f2.done ? f2.out : ln(t1, p + ".norm2");
        Tensor hdn = linear(n2, p + ".linear1", ACT_GELU);
        tfree(n2);
        pre = LnFuse();
        if (l < 8) { pre.prefix = "ft_layers." + std::to_string(l + 1) + ".norm1"; pre.add2 = pos; pre.add2_rows = L; }
        else pre.prefix = "idx_pred_layer.0";
        t = linear(hdn, p + ".linear2", ACT_NONE, &t1, &pre);
        tfree(hdn); tfree(t1);
    }
    Tensor tn = pre.done ? pre.out : ln(t, "idx_pred_layer.0");
    tfree(t);
    Tensor logits = linear(tn, "idx_pred_layer.1");
    tfree(tn);
    Tensor quant = talloc(nbt, 16, 16, 256, adt_);
    int* idx = (int*)ar_->alloc((size_t)nbt * L * sizeof(int));
    if (!ar_->dry()) {
        const int* forced = nullptr;
        auto it = forced_.find("codes");
        if (it != forced_.end() && it->second.p) forced = (const int*)it->second.p + (size_t)frame * L;   // per-clip path only
        argmax_gather(logits.f(), nbt * L, 1024, warr("quantize.embedding.weight"), 256, forced, idx, quant.p, quant.dt, s_, status_);
        launches_ += 1;
        if (capture_) {
            Cap& cl = cap_["logits"];
            Cap& cc = cap_["codes"];
            if (cl.p && (size_t)(frame + 1) * L * 1024 * 4 <= cl.bytes)
                CUDA_CHECK(cudaMemcpyAsync((float*)cl.p + (size_t)frame * L * 1024, logits.p, (size_t)L * 1024 * 4,
                                           cudaMemcpyDeviceToDevice, s_));
            if (cc.p && (size_t)(frame + 1) * L * 4 <= cc.bytes)
                CUDA_CHECK(cudaMemcpyAsync((int*)cc.p + (size_t)frame * L, idx, (size_t)L * 4, cudaMemcpyDeviceToDevice, s_));
        }
    }
    ar_->free(idx);
    tfree(logits);
    return quant;
}

// Fuse_sft_block.forward (keep_arch.py:465-472)
Tensor Engine::cft(const Tensor& enc, const Tensor& dec, const std::string& p) {
    Tensor f = res_block(enc, p + ".encode_enc", &dec);
    ConvOpt o1;
    o1.pad(1); o1.act = ACT_LRELU02;
    ConvOpt o2;
    o2.pad(1);
    Tensor s1 = conv(f, p + ".scale.0", o1);
    Tensor sc = conv(s1, p + ".scale.2", o2);
    tfree(s1);
    Tensor t1 = conv(f, p + ".shift.0", o1);
    Tensor sh = conv(t1, p + ".shift.2", o2);
    tfree(t1); tfree(f);
    Tensor out = talloc(dec.n, dec.h, dec.w, dec.c, dec.dt);
    if (!ar_->dry()) { cft_combine(dec.p, dec.dt, sc.p, sh.p, sc.dt, 1.0f, out.p, out.dt, out.numel(), s_); launches_ += 1; }
    tfree(sc); tfree(sh);
    return out;
}

// CrossFrameFusionLayer.forward, residual=True (keep_arch.py:519-541); 4 heads x 256
Tensor Engine::cfa(const Tensor& cur, const Tensor& prev, const std::string& p) {
    KEEP_CHECK(cur.dt == F32 && prev.dt == F32, "cfa expects fp32 feature maps");
    const int L = cur.h * cur.w, C = cur.c, heads = 4, dh = 256, inner = heads * dh;
    const int nbt = cur.n;   // clips in lockstep (1 on the per-clip path)
    KEEP_CHECK(prev.n == nbt, "cfa: current / previous feature batch mismatch");
    Tensor x = cur, pv = prev;
    x.n = 1; x.h = nbt * L; x.w = 1;
    pv.n = 1; pv.h = nbt * L; pv.w = 1;
    Tensor q = linear(x, p + ".attn.to_q"), k = linear(pv, p + ".attn.to_k"), v = linear(pv, p + ".attn.to_v");
    const long long bs = (long long)L * inner;
    static const int tc_min_L = getenv("KEEP_MHA_TC_MIN_L") ? atoi(getenv("KEEP_MHA_TC_MIN_L")) : 256;   // (1024: only the 32^2 block; both measured equal in time, profiles/r2_experiments.md)
    Tensor o = ((flags_ & KEEP_FLAG_TCGEN05) && nbt == 1 && L >= tc_min_L)
                   ? mha_tc(q.f(), k.f(), v.f(), L, L, heads, dh, 1.0f / sqrtf((float)dh))
                   : mha(q.f(), inner, bs, k.f(), inner, bs, v.f(), inner, bs, nbt, L, L, heads, dh, 1.0f / sqrtf((float)dh));
    tfree(q); tfree(k); tfree(v);
    Tensor y = linear(o, p + ".attn.to_out.0");
    tfree(o);
    Tensor x1 = ln(y, p + ".norm1", &x);
    tfree(y);
    Tensor pr = linear(x1, p + ".ff.net.0.proj");
    Tensor gg = talloc(1, nbt * L, 1, 4 * C, F32);
    if (!ar_->dry()) { geglu(pr.f(), gg.f(), nbt * L, 4 * C, s_); launches_ += 1; }
    tfree(pr);
    Tensor y2 = linear(gg, p + ".ff.net.2");
    tfree(gg);
    Tensor x2 = ln(y2, p + ".norm2", &x1);
    tfree(y2); tfree(x1);
    x2.n = cur.n; x2.h = cur.h; x2.w = cur.w;
    return x2;
}

// Generator programme with CFT / CFA hooks (keep_arch.py:1101-1125)
Tensor Engine::generator(const Tensor& quant, int frame, Tensor taps[6], Tensor cfa_prev[6]) {
    Tensor x = quant;
    bool own = false;
    Aff pend;
    // KEEP_GEN_FAST_FROM=j (experiment, default off): generator blocks >= j run with plain fp16 operands (1 MMA pass instead
    // of 3) in the split-precision mode -- they are the last layers before the pixels, so their rounding error is not
    // amplified by later normalisations; 20 = the 512^2 level, 17 = + the 256^2 level
    static const int fast_from = getenv("KEEP_GEN_FAST_FROM") ? atoi(getenv("KEEP_GEN_FAST_FROM")) : 99;
    wide_scope_ = (flags_ & KEEP_FLAG_TC_WIDE) != 0;
    for (int j = 0; j < 25; ++j) {
        pass_override_ = j >= fast_from ? 1 : 0;
        const std::string bp = "generator.blocks." + std::to_string(j);
        const std::string kind = kGenProg[j];
        // the next block normalises this block's output -- unless a CFT / CFA hook replaces it first (then the hook's output
        // is what gets normalised, by the stand-alone statistics kernel)
        const std::string next = j + 1 < 25 ? kGenProg[j + 1] : "";
        bool next_gn = next == "res" || next == "attn" || next == "norm";
        for (int k = 0; k < 6; ++k)
            if (kFuseGen[k] == j && (cft_on_[k] || (cfa_on_[k] && frame > 0))) next_gn = false;
        const std::string nbp = "generator.blocks." + std::to_string(j + 1);
        const std::string next_norm = !next_gn ? "" : (next == "res" ? nbp + ".norm1" : (next == "attn" ? nbp + ".norm" : nbp));
        Tensor y;
        if (kind == "conv") {
            ConvOpt o;
            o.pad(1); o.want_stats = next_gn; o.stats_norm = next_norm;
            if (j == 24) { o.pre = &pend; o.out_dt = F32; }
            y = conv(x, bp, o);
            if (j == 24) afree(pend);
        } else if (kind == "res") {
            y = res_block(x, bp, nullptr, next_gn, next_norm);
        } else if (kind == "attn") {
            y = attn_block(x, bp, next_gn, next_norm);
        } else if (kind == "up") {  // vqgan_arch.py:148-152: nearest x2 fused into the conv's gather
            ConvOpt o;
            o.pad(1); o.up = 2; o.want_stats = next_gn; o.stats_norm = next_norm;
            y = conv(x, bp + ".conv", o);
        } else {
            pend = gn(x, bp);
            continue;
        }
        if (own) tfree(x);
        x = y;
        own = true;
        int ti = -1;
        for (int k = 0; k < 6; ++k) if (kFuseGen[k] == j) ti = k;
        if (ti < 0) continue;
        const std::string sz = std::to_string(kFuseSize[ti]);
        if (cft_on_[ti]) {   // keep_arch.py:1104-1108
            Tensor enc = taps[ti];   // (T * nb, s, s, C), frame-major: the nb maps of this frame (nb = 1 on the per-clip path)
            enc.n = x.n;
            enc.p = (char*)enc.p + (size_t)frame * x.n * enc.h * enc.w * enc.c * dtype_size(enc.dt);
            Tensor z = cft(enc, x, "cft." + sz);
            tfree(x);
            x = z;
        }
        if (cfa_on_[ti]) {   // keep_arch.py:1110-1121: the previous frame's *fused* feature is the key / value source
            if (frame > 0) {
                Tensor z2 = cfa(x, cfa_prev[ti], "cfa." + sz);
                tfree(x);
                x = z2;
            }
            if (!ar_->dry())
                CUDA_CHECK(cudaMemcpyAsync(cfa_prev[ti].p, x.p, x.bytes(), cudaMemcpyDeviceToDevice, s_));
        }
    }
    pass_override_ = 0;
    wide_scope_ = false;
    return x;   // (1,512,512,3) fp32
}

// =============================================================================================
// full forward for one clip (b = 1)
// =============================================================================================
void Engine::forward_clip(const float* x_dev, int T, void* out_dev, int out_dtype) {
    const int HW = 512 * 512;
    const bool dry = ar_->dry();
    // ---- persistent per-clip tensors
    Tensor flows = talloc(T - 1, 512, 512, 2, F32);
    Tensor taps[6], cfa_prev[6];
    for (int k = 0; k < 6; ++k)
        if (cft_on_[k]) taps[k] = talloc(T, kFuseSize[k], kFuseSize[k], kFuseCh[k], adt_);
    Tensor z_codes = talloc(T, 16, 16, 256, F32);
    for (int k = 0; k < 6; ++k)
        if (cfa_on_[k]) cfa_prev[k] = talloc(1, kFuseSize[k], kFuseSize[k], kFuseCh[k], F32);

    // ---- optical flow (batched over pairs)
    auto ff = forced_.find("flows");
    bool flows_async = false;
    static const bool skip_flow = getenv("KEEP_DEBUG_SKIP_FLOW") != nullptr;   // timing experiments only: zero flows, no GMFlow
    if (!dry && ff != forced_.end() && ff->second.p) {
        KEEP_CHECK(ff->second.bytes == flows.bytes(), "forced flows have the wrong size");
        CUDA_CHECK(cudaMemcpyAsync(flows.p, ff->second.p, flows.bytes(), cudaMemcpyDeviceToDevice, s_));
    } else if (!dry && skip_flow) {
        CUDA_CHECK(cudaMemsetAsync(flows.p, 0, flows.bytes(), s_));
    } else {
        // fork: GMFlow for all pairs runs on the side stream / side arena, overlapping the LQ encoder, the gain
        // estimator and the serial per-frame chain (frame i only needs the flow of pair i-1).
        // GMFlow inline on the main stream: KEEP_NO_SIDE (debug / timeline), and the per-launch profiling pass -- its CUDA events
        // must time every kernel ALONE (with the side branch running, a full-grid main-stream kernel shares the SMs with
        // GMFlow's persistent CTAs and its event-to-event time is not the kernel's own: 84 us vs 62 us for the dominant conv shape)
        static const bool env_no_side = getenv("KEEP_NO_SIDE") != nullptr;
        const bool no_side = env_no_side || profile_;
        if (!dry && !no_side) {
            if (!side_) {
                int lo = 0, hi = 0;
                CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));       // lo = least urgent
                CUDA_CHECK(cudaStreamCreateWithPriority(&side_, cudaStreamNonBlocking, lo));
                CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
            }
            while ((int)ev_flow_.size() < (T - 1 + flow_chunk() - 1) / flow_chunk()) {
                cudaEvent_t e;
                CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ev_flow_.push_back(e);
            }
            CUDA_CHECK(cudaEventRecord(ev_fork_, s_main_));
            CUDA_CHECK(cudaStreamWaitEvent(side_, ev_fork_, 0));
            s_ = side_;
        }
        ar_ = &arena2_;
        gmflow(x_dev, T, flows.f());
        ar_ = &arena_;
        s_ = s_main_;
        flows_async = !dry && !no_side;
    }
    if (capture_ && flows_async) {   // debug capture wants the complete flows now
        CUDA_CHECK(cudaStreamWaitEvent(s_main_, ev_flow_[(T - 2) / flow_chunk()], 0));
        flows_async = false;
    }
    capture("flows", flows.p, flows.bytes());

    // While the GMFlow branch holds side_sms_ SMs with long-lived persistent CTAs, a full-width main-stream grid would have
    // its last CTAs queue behind them; KEEP_MAIN_SMS caps main-stream persistent grids for the LQ encoder and the first
    // KEEP_MAIN_CAP_FRAMES frames of the recurrence (the stretch GMFlow overlaps).
    static const int env_main_sms = getenv("KEEP_MAIN_SMS") ? atoi(getenv("KEEP_MAIN_SMS")) : 0;
    static const int env_cap_frames = getenv("KEEP_MAIN_CAP_FRAMES") ? atoi(getenv("KEEP_MAIN_CAP_FRAMES")) : 8;
    main_cap_ = (flows_async && env_main_sms > 0) ? std::min(env_main_sms, num_sms_) : num_sms_;
    // ---- LQ encoder, batched over frames in chunks (keep_arch.py:1034-1037)
    const int chunk = lq_chunk();
    for (int f0 = 0; f0 < T; f0 += chunk) {
        const int nf = std::min(chunk, T - f0);
        Tensor img = talloc(nf, 512, 512, 3, adt_);
        if (!dry) { nchw_to_nhwc(x_dev + (size_t)f0 * 3 * HW, img.p, img.dt, nf, 3, 512, 512, 0, s_); launches_ += 1; }
        auto tap = [&](int i, const Tensor& t) {
            int ti = -1;
            for (int k = 0; k < 6; ++k) if (kFuseEnc[k] == i && cft_on_[k]) ti = k;
            if (ti < 0 || dry) return;
            const size_t per = (size_t)t.h * t.w * t.c * dtype_size(t.dt);
            CUDA_CHECK(cudaMemcpyAsync((char*)taps[ti].p + (size_t)f0 * per, t.p, per * nf, cudaMemcpyDeviceToDevice, s_));
        };
        Tensor z = encoder(img, "encoder", tap);
        if (!dry)
            CUDA_CHECK(cudaMemcpyAsync(z_codes.f() + (size_t)f0 * 256 * 256, z.p, z.bytes(), cudaMemcpyDeviceToDevice, s_));
        tfree(z);
        tfree(img);
    }
    capture("z_codes", z_codes.p, z_codes.bytes());

    // ---- Kalman gains (batched over T)
    Tensor gains = kalman_gains(z_codes, T);
    capture("gains", gains.p, gains.bytes());
    if (capture_ && !dry) {
        for (const char* nm : {"logits", "codes"}) {
            Cap& c = cap_[nm];
            const size_t need = std::string(nm) == "logits" ? (size_t)T * 256 * 1024 * 4 : (size_t)T * 256 * 4;
            if (c.bytes < need) { cudaFree(c.p); CUDA_CHECK(cudaMalloc(&c.p, need)); c.bytes = need; }
        }
    }

    // ---- serial per-frame recurrence (keep_arch.py:1062-1128)
    Tensor prev_out;  // (1,512,512,3) fp32 NHWC
    auto fp = forced_.find("prev");
    const bool force_prev = dry || (fp != forced_.end() && fp->second.p);
    if (!dry) {   // forced buffers are indexed by frame below: refuse short ones instead of reading out of bounds
        if (fp != forced_.end() && fp->second.p)
            KEEP_CHECK(fp->second.bytes >= (size_t)(T - 1) * 3 * HW * sizeof(float), "forced 'prev' holds fewer than T-1 = %d frames", T - 1);
        auto fc = forced_.find("codes");
        if (fc != forced_.end() && fc->second.p)
            KEEP_CHECK(fc->second.bytes >= (size_t)T * 256 * sizeof(int), "forced 'codes' holds fewer than T = %d frames of 256 indices", T);
    }
    for (int i = 0; i < T; ++i) {
        if (i >= env_cap_frames) main_cap_ = num_sms_;
        Tensor z_hat;
        bool own_z = false;
        if (i == 0) {
            z_hat = z_codes;
            z_hat.n = 1;
        } else {
            Tensor src = prev_out;
            bool own_src = false;
            if (force_prev) {
                src = talloc(1, 512, 512, 3, F32);
                own_src = true;
                if (!dry) { nchw_to_nhwc((const float*)fp->second.p + (size_t)(i - 1) * 3 * HW, src.p, F32, 1, 3, 512, 512, 0, s_); launches_ += 1; }
            }
            if (flows_async) CUDA_CHECK(cudaStreamWaitEvent(s_main_, ev_flow_[(i - 1) / flow_chunk()], 0));   // flow of pair i-1 is ready
            Tensor warped = talloc(1, 512, 512, 3, adt_);
            if (!dry) {
                flow_warp(src.p, src.dt, flows.f() + (size_t)(i - 1) * HW * 2, warped.p, warped.dt, 1, 512, 512, 3, s_, status_);
                launches_ += 1;
            }
            if (own_src) tfree(src);
            Tensor zp = encoder(warped, "hq_encoder", nullptr);
            tfree(warped);
            z_hat = talloc(1, 16, 16, 256, F32);
            own_z = true;
            if (!dry) {
                kalman_update(z_codes.f() + (size_t)i * 256 * 256, zp.f(), gains.f() + (size_t)i * 256, z_hat.f(), 256, 256, s_, status_);
                launches_ += 1;
            }
            tfree(zp);
        }
        Tensor zh = z_hat;
        if (i == 0) zh.p = z_codes.p;
        Tensor quant = code_transformer(zh, i);
        if (own_z) tfree(z_hat);
        Tensor img = generator(quant, i, taps, cfa_prev);
        tfree(quant);
        if (!dry) {
            if (out_dtype == KEEP_OUT_U8_BGR)
                nhwc_to_u8bgr(img.p, img.dt, (unsigned char*)out_dev + (size_t)i * 3 * HW, 1, 512, 512, s_, status_);
            else
                nhwc_to_nchw(img.p, img.dt, (char*)out_dev + (size_t)i * 3 * HW * (out_dtype == KEEP_OUT_F16 ? 2 : 4),
                             out_dtype == KEEP_OUT_F16 ? F16 : F32, 1, 3, 512, 512, s_, status_);
            launches_ += 1;
        }
        if (prev_out.p) tfree(prev_out);
        prev_out = img;
    }
    main_cap_ = num_sms_;
    if (flows_async) CUDA_CHECK(cudaStreamWaitEvent(s_main_, ev_flow_[(T - 2) / flow_chunk()], 0));   // join the side branch
    if (prev_out.p) tfree(prev_out);
    tfree(gains);
    for (int k = 5; k >= 0; --k) if (cfa_on_[k]) tfree(cfa_prev[k]);
    tfree(z_codes);
    for (int k = 0; k < 6; ++k) if (cft_on_[k]) tfree(taps[k]);
    tfree(flows);
}

// =============================================================================================
// lockstep forward for a group of nb clips (KEEP_FLAG_BATCH_CLIPS)
// =============================================================================================
// Clips are independent (keep_processor.py:263-270) but each one is a serial chain of ~500 small kernels per frame that
// cannot fill 148 SMs (16^2 .. 64^2 maps: 2-32 M tiles per layer).  Here nb clips walk the recurrence together: frame i of
// every clip goes through ONE hq_encoder / code-transformer / generator pass with batch nb, so each launch carries nb times
// the work (fewer K-splits per layer, the fixed per-kernel cost paid once per nb frames).  Per-clip tensors that the
// recurrence slices by frame (encoder taps, z_codes, gains) are stored frame-major -- slot i*nb + c -- so the nb maps
// of one frame index are one contiguous batch.  GMFlow, the LQ encoder and the gain estimator are already batched over
// frames and run clip by clip, as on the per-clip path.  No debug forcing / capture on this path.
void Engine::forward_clips(const float* x_dev, int nb, int T, void* out_dev, int out_dtype) {
    const int HW = 512 * 512;
    const bool dry = ar_->dry();
    const size_t per_clip_in = (size_t)T * 3 * HW;
    const size_t osz = out_dtype == KEEP_OUT_F16 ? 2 : (out_dtype == KEEP_OUT_U8_BGR ? 1 : 4);
    std::vector<Tensor> flows(nb);
    for (int c = 0; c < nb; ++c) flows[c] = talloc(T - 1, 512, 512, 2, F32);
    Tensor taps[6], cfa_prev[6];
    for (int k = 0; k < 6; ++k)
        if (cft_on_[k]) taps[k] = talloc(T * nb, kFuseSize[k], kFuseSize[k], kFuseCh[k], adt_);
    Tensor z_all = talloc(T * nb, 16, 16, 256, F32);
    Tensor gains_all = talloc(T * nb, 16, 16, 1, F32);
    for (int k = 0; k < 6; ++k)
        if (cfa_on_[k]) cfa_prev[k] = talloc(nb, kFuseSize[k], kFuseSize[k], kFuseCh[k], F32);
    // rows of `per` bytes, one per frame of clip c: contiguous (T, per) <-> frame-major slots (i*nb + c)
    auto scatter_frames = [&](void* dst_all, const void* src, size_t per, int f0, int nf, int c) {
        CUDA_CHECK(cudaMemcpy2DAsync((char*)dst_all + ((size_t)f0 * nb + c) * per, (size_t)nb * per, src, per, per, nf,
                                     cudaMemcpyDeviceToDevice, s_));
    };

    // ---- optical flow of every clip on the side stream (low priority), overlapping everything below
    static const bool env_no_side = getenv("KEEP_NO_SIDE") != nullptr;
    const bool no_side = env_no_side || profile_;   // (profiling pass: every kernel timed alone, see forward_clip)
    static const bool skip_flow = getenv("KEEP_DEBUG_SKIP_FLOW") != nullptr;   // timing experiments only: zero flows, no GMFlow
    const int nchunk = (T - 1 + flow_chunk() - 1) / flow_chunk();
    const bool flows_async = !dry && !no_side && !skip_flow;
    if (flows_async) {
        if (!side_) {
            int lo = 0, hi = 0;
            CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CUDA_CHECK(cudaStreamCreateWithPriority(&side_, cudaStreamNonBlocking, lo));
            CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
        }
        while ((int)ev_flow_.size() < nb * nchunk) {
            cudaEvent_t e;
            CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ev_flow_.push_back(e);
        }
        CUDA_CHECK(cudaEventRecord(ev_fork_, s_main_));
        CUDA_CHECK(cudaStreamWaitEvent(side_, ev_fork_, 0));
        s_ = side_;
    }
    ar_ = &arena2_;
    // chunk-major order: frame i of the lockstep group needs pair i-1 of EVERY clip, so chunk k of all clips goes before
    // chunk k+1 of any (clip-major order would hold frame 1 back until nearly all of the group's GMFlow work is done)
    if (!dry && skip_flow) {
        for (int c = 0; c < nb; ++c) CUDA_CHECK(cudaMemsetAsync(flows[c].p, 0, flows[c].bytes(), s_));
    } else {
        for (int p0 = 0; p0 < T - 1; p0 += flow_chunk())
            for (int c = 0; c < nb; ++c) {
                ev_flow_base_ = c * nchunk;
                gmflow(x_dev + (size_t)c * per_clip_in, T, flows[c].f(), p0, p0 + flow_chunk());
            }
    }
    ev_flow_base_ = 0;
    ar_ = &arena_;
    s_ = s_main_;
    main_cap_ = num_sms_;

    // ---- LQ encoder (frames of one clip per pass, keep_arch.py:1034-1037) and Kalman gains, clip by clip
    const int chunk = lq_chunk();
    for (int c = 0; c < nb; ++c) {
        const float* xc = x_dev + (size_t)c * per_clip_in;
        Tensor zc = talloc(T, 16, 16, 256, F32);
        for (int f0 = 0; f0 < T; f0 += chunk) {
            const int nf = std::min(chunk, T - f0);
            Tensor img = talloc(nf, 512, 512, 3, adt_);
            if (!dry) { nchw_to_nhwc(xc + (size_t)f0 * 3 * HW, img.p, img.dt, nf, 3, 512, 512, 0, s_); launches_ += 1; }
            auto tap = [&](int i, const Tensor& t) {
                int ti = -1;
                for (int k = 0; k < 6; ++k) if (kFuseEnc[k] == i && cft_on_[k]) ti = k;
                if (ti < 0 || dry) return;
                scatter_frames(taps[ti].p, t.p, (size_t)t.h * t.w * t.c * dtype_size(t.dt), f0, nf, c);
            };
            Tensor z = encoder(img, "encoder", tap);
            if (!dry) {
                CUDA_CHECK(cudaMemcpyAsync(zc.f() + (size_t)f0 * 256 * 256, z.p, z.bytes(), cudaMemcpyDeviceToDevice, s_));
                scatter_frames(z_all.p, z.p, (size_t)256 * 256 * sizeof(float), f0, nf, c);
            }
            tfree(z);
            tfree(img);
        }
        Tensor g = kalman_gains(zc, T);   // (T, 16, 16, 1)
        if (!dry) scatter_frames(gains_all.p, g.p, (size_t)256 * sizeof(float), 0, T, c);
        tfree(g);
        tfree(zc);
    }

    // ---- the recurrence, all clips in lockstep (keep_arch.py:1062-1128)
    Tensor prev_out;   // (nb, 512, 512, 3) fp32 NHWC
    for (int i = 0; i < T; ++i) {
        Tensor z_hat;
        bool own_z = false;
        if (i == 0) {
            z_hat = z_all;   // slots 0 .. nb-1 = frame 0 of every clip
            z_hat.n = nb;
        } else {
            Tensor warped = talloc(nb, 512, 512, 3, adt_);
            for (int c = 0; c < nb; ++c) {
                if (flows_async) CUDA_CHECK(cudaStreamWaitEvent(s_main_, ev_flow_[c * nchunk + (i - 1) / flow_chunk()], 0));
                if (!dry) {
                    flow_warp((const char*)prev_out.p + (size_t)c * HW * 3 * dtype_size(prev_out.dt), prev_out.dt,
                              flows[c].f() + (size_t)(i - 1) * HW * 2, (char*)warped.p + (size_t)c * HW * 3 * dtype_size(warped.dt),
                              warped.dt, 1, 512, 512, 3, s_, status_);
                    launches_ += 1;
                }
            }
            Tensor zp = encoder(warped, "hq_encoder", nullptr);   // (nb, 16, 16, 256)
            tfree(warped);
            z_hat = talloc(nb, 16, 16, 256, F32);
            own_z = true;
            if (!dry) {
                kalman_update(z_all.f() + (size_t)i * nb * 256 * 256, zp.f(), gains_all.f() + (size_t)i * nb * 256, z_hat.f(), nb * 256,
                              256, s_, status_);
                launches_ += 1;
            }
            tfree(zp);
        }
        Tensor quant = code_transformer(z_hat, i);
        if (own_z) tfree(z_hat);
        Tensor img = generator(quant, i, taps, cfa_prev);   // (nb, 512, 512, 3)
        tfree(quant);
        if (!dry) {
            for (int c = 0; c < nb; ++c) {
                const char* src = (const char*)img.p + (size_t)c * HW * 3 * dtype_size(img.dt);
                char* dst = (char*)out_dev + ((size_t)c * T + i) * 3 * HW * osz;
                if (out_dtype == KEEP_OUT_U8_BGR) nhwc_to_u8bgr(src, img.dt, (unsigned char*)dst, 1, 512, 512, s_, status_);
                else nhwc_to_nchw(src, img.dt, dst, out_dtype == KEEP_OUT_F16 ? F16 : F32, 1, 3, 512, 512, s_, status_);
                launches_ += 1;
            }
        }
        if (prev_out.p) tfree(prev_out);
        prev_out = img;
    }
    if (flows_async)
        for (int c = 0; c < nb; ++c) CUDA_CHECK(cudaStreamWaitEvent(s_main_, ev_flow_[c * nchunk + nchunk - 1], 0));   // join the side branch
    if (prev_out.p) tfree(prev_out);
    for (int k = 5; k >= 0; --k) if (cfa_on_[k]) tfree(cfa_prev[k]);
    tfree(gains_all);
    tfree(z_all);
    for (int k = 0; k < 6; ++k) if (cft_on_[k]) tfree(taps[k]);
    for (int c = 0; c < nb; ++c) tfree(flows[c]);
}

void Engine::plan_dump(int nb, int T, const char* path) {
    KEEP_CHECK(nb >= 1 && nb <= 8 && T >= 2 && T <= 100, "keep_plan_dump: need 1 <= clips <= 8 and 2 <= T <= 100");
    std::vector<std::string> lines;
    plan_ = &lines;
    try {
        begin(nullptr, 0, nullptr, true);
        if (nb == 1) forward_clip(nullptr, T, nullptr, KEEP_OUT_F32);
        else forward_clips(nullptr, nb, T, nullptr, KEEP_OUT_F32);
    } catch (...) {
        plan_ = nullptr;
        throw;
    }
    plan_ = nullptr;
    FILE* f = fopen(path, "w");
    KEEP_CHECK(f, "cannot open %s", path);
    for (auto& l : lines) fprintf(f, "%s\n", l.c_str());
    fclose(f);
}

// workspace plan of a lockstep group (dry run), cached per (nb, T)
size_t Engine::plan_clips(int nb, int T) {
    const int key = nb * 1000 + T;
    auto it = ws_cache_.find(key);
    if (it != ws_cache_.end()) return it->second;
    begin(nullptr, 0, nullptr, true);
    forward_clips(nullptr, nb, T, nullptr, KEEP_OUT_F32);
    const size_t side = (arena2_.peak() + 4095) & ~(size_t)4095;
    const size_t need = ((arena_.peak() + 4095) & ~(size_t)4095) + side + 4096;
    ws_cache_[key] = need;
    side_cache_[key] = side;
    return need;
}

// keep_forward with KEEP_FLAG_BATCH_CLIPS and b > 1: groups of up to batch_max_ clips in lockstep, a trailing single clip on
// the per-clip path.  Engine-owned workspace; CUDA-graph replay per (group size, T) like the per-clip path.
void Engine::forward_batched(const float* x_dev, int b, int T, void* out_dev, int out_dtype, cudaStream_t s) {
    const size_t per_clip = (size_t)T * 3 * 512 * 512;
    const size_t osz = out_dtype == KEEP_OUT_F16 ? 2 : (out_dtype == KEEP_OUT_U8_BGR ? 1 : 4);
    for (int b0 = 0; b0 < b;) {
        const int g = std::min(batch_max_, b - b0);
        const float* xin = x_dev + (size_t)b0 * per_clip;
        char* xout = (char*)out_dev + (size_t)b0 * per_clip * osz;
        if (g == 1) {   // odd clip out: the ordinary per-clip call
            forward(xin, 1, T, xout, out_dtype, nullptr, 0, s);
            b0 += 1;
            continue;
        }
        const int key = g * 1000 + T;
        const size_t need = plan_clips(g, T);
        if (own_ws_bytes_ < need) {
            CUDA_CHECK(cudaStreamSynchronize(s));
            cudaFree(own_ws_);
            own_ws_ = nullptr; own_ws_bytes_ = 0;
            for (auto& gr : graphs_) cudaGraphExecDestroy(gr.second.exec);
            graphs_.clear();
            CUDA_CHECK(cudaMalloc(&own_ws_, need));
            own_ws_bytes_ = need;
        }
        side_bytes_ = side_cache_[key];
        void* ws = own_ws_;
        const size_t ws_bytes = own_ws_bytes_;
        const bool want_graph = (flags_ & KEEP_FLAG_CUDA_GRAPH) != 0;
        if (want_graph && (eager_runs_[key] >= 1 || prepacked_)) {
            const size_t in_bytes = (size_t)g * per_clip * 4;
            if (gx_bytes_ < in_bytes || gout_bytes_ < in_bytes) {
                CUDA_CHECK(cudaStreamSynchronize(s));
                cudaFree(gx_); cudaFree(gout_);
                gx_ = nullptr; gout_ = nullptr;
                for (auto& gr : graphs_) cudaGraphExecDestroy(gr.second.exec);
                graphs_.clear();
                CUDA_CHECK(cudaMalloc((void**)&gx_, in_bytes));
                CUDA_CHECK(cudaMalloc(&gout_, in_bytes));
                gx_bytes_ = gout_bytes_ = in_bytes;
            }
            if (!gs_) {
                int lo = 0, hi = 0;
                CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                CUDA_CHECK(cudaStreamCreateWithPriority(&gs_, cudaStreamNonBlocking, hi));
                CUDA_CHECK(cudaEventCreateWithFlags(&ev_in_, cudaEventDisableTiming));
                CUDA_CHECK(cudaEventCreateWithFlags(&ev_out_, cudaEventDisableTiming));
            }
            ClipGraph& gr = graphs_[key];
            if (gr.exec && (gr.ws != ws || gr.out_dtype != out_dtype)) { cudaGraphExecDestroy(gr.exec); gr.exec = nullptr; }
            CUDA_CHECK(cudaEventRecord(ev_in_, s));
            CUDA_CHECK(cudaStreamWaitEvent(gs_, ev_in_, 0));
            if (!gr.exec) {
                cudaGraph_t graph = nullptr;
                const long long l0 = launches_;
                CUDA_CHECK(cudaStreamBeginCapture(gs_, cudaStreamCaptureModeThreadLocal));
                try {
                    begin(ws, ws_bytes, gs_, false);
                    forward_clips(gx_, g, T, gout_, out_dtype);
                } catch (...) {
                    cudaStreamEndCapture(gs_, &graph);
                    if (graph) cudaGraphDestroy(graph);
                    throw;
                }
                launches_per_clip_[key] = launches_ - l0;
                launches_ = l0;
                CUDA_CHECK(cudaStreamEndCapture(gs_, &graph));
                cudaError_t e = cudaGraphInstantiate(&gr.exec, graph, 0);
                cudaGraphDestroy(graph);
                CUDA_CHECK(e);
                gr.ws = ws; gr.out_dtype = out_dtype;
            }
            CUDA_CHECK(cudaMemcpyAsync(gx_, xin, (size_t)g * per_clip * 4, cudaMemcpyDeviceToDevice, gs_));
            CUDA_CHECK(cudaGraphLaunch(gr.exec, gs_));
            CUDA_CHECK(cudaMemcpyAsync(xout, gout_, (size_t)g * per_clip * osz, cudaMemcpyDeviceToDevice, gs_));
            CUDA_CHECK(cudaEventRecord(ev_out_, gs_));
            CUDA_CHECK(cudaStreamWaitEvent(s, ev_out_, 0));
            launches_ += launches_per_clip_[key];
        } else {
            const long long l0 = launches_;
            begin(ws, ws_bytes, s, false);
            forward_clips(xin, g, T, xout, out_dtype);
            launches_per_clip_[key] = launches_ - l0;
            eager_runs_[key] += 1;
        }
        b0 += g;
    }
}

size_t Engine::workspace_bytes(int b, int T) {
    KEEP_CHECK(b >= 1 && T >= 2 && T <= 100, "keep_workspace_bytes: need b >= 1 and 2 <= T <= 100 (got b=%d T=%d)", b, T);
    if ((flags_ & KEEP_FLAG_BATCH_CLIPS) && b > 1 && batch_max_ > 1)   // engine-owned workspace of a lockstep group (informational)
        return std::max(plan_clips(std::min(b, batch_max_), T), workspace_bytes(1, T));
    auto it = ws_cache_.find(T);
    if (it != ws_cache_.end()) return it->second;
    begin(nullptr, 0, nullptr, true);
    forward_clip(nullptr, T, nullptr, KEEP_OUT_F32);
    const size_t side = (arena2_.peak() + 4095) & ~(size_t)4095;
    const size_t need = ((arena_.peak() + 4095) & ~(size_t)4095) + side + 4096;
    ws_cache_[T] = need;
    side_cache_[T] = side;
    return need;
}

void Engine::forward(const float* x_dev, int b, int T, void* out_dev, int out_dtype, void* ws, size_t ws_bytes, cudaStream_t s) {
    KEEP_CHECK(!dry_only_, "keep_forward: engine was created with KEEP_FLAG_PLAN_ONLY (no device)");
    KEEP_CHECK(x_dev && out_dev, "keep_forward: null tensor");
    KEEP_CHECK(b >= 1 && T >= 2 && T <= 100, "keep_forward: need b >= 1 and 2 <= T <= 100 (got b=%d T=%d)", b, T);
    KEEP_CHECK(out_dtype == KEEP_OUT_F32 || out_dtype == KEEP_OUT_F16 || out_dtype == KEEP_OUT_U8_BGR, "keep_forward: bad out_dtype %d", out_dtype);
    CUDA_CHECK(cudaSetDevice(device_));
    // Engine-owned buffers (workspace, staging, lazily packed weight panels) are ordered by the stream of the call that
    // touches them; a caller that switches streams between calls gets the dependency from this event instead of a race.
    CUDA_CHECK(cudaStreamWaitEvent(s, ev_last_, 0));
    struct Tail { cudaEvent_t e; cudaStream_t s; ~Tail() { cudaEventRecord(e, s); } } tail{ev_last_, s};
    if ((flags_ & KEEP_FLAG_BATCH_CLIPS) && b > 1 && batch_max_ > 1 && !ws && !capture_ && !profile_) {
        bool forcing_b = false;
        for (auto& kv : forced_) forcing_b = forcing_b || kv.second.p != nullptr;
        if (!forcing_b) { forward_batched(x_dev, b, T, out_dev, out_dtype, s); return; }
    }
    const size_t need = workspace_bytes(1, T);
    side_bytes_ = side_cache_[T];
    if (!ws) {
        if (own_ws_bytes_ < need) {
            CUDA_CHECK(cudaStreamSynchronize(s));
            cudaFree(own_ws_);
            own_ws_ = nullptr; own_ws_bytes_ = 0;
            for (auto& g : graphs_) cudaGraphExecDestroy(g.second.exec);
            graphs_.clear();
            CUDA_CHECK(cudaMalloc(&own_ws_, need));
            own_ws_bytes_ = need;
        }
        ws = own_ws_; ws_bytes = own_ws_bytes_;
    }
    KEEP_CHECK(ws_bytes >= need, "keep_forward: workspace too small (%zu < %zu)", ws_bytes, need);
    KEEP_CHECK(((uintptr_t)ws & 255) == 0, "keep_forward: workspace must be 256-byte aligned");
    const size_t per_clip = (size_t)T * 3 * 512 * 512;
    const size_t osz = out_dtype == KEEP_OUT_F16 ? 2 : (out_dtype == KEEP_OUT_U8_BGR ? 1 : 4);
    bool forcing = false;
    for (auto& kv : forced_) forcing = forcing || kv.second.p != nullptr;
    const bool want_graph = (flags_ & KEEP_FLAG_CUDA_GRAPH) && !capture_ && !profile_ && !forcing && ws == own_ws_;
    for (int bi = 0; bi < b; ++bi) {   // clips are independent (keep_processor.py:263-270)
        const float* xin = x_dev + bi * per_clip;
        char* xout = (char*)out_dev + bi * per_clip * osz;
        if (want_graph && (eager_runs_[T] >= 1 || prepacked_)) {   // (prepacked weights: nothing is allocated inside a forward, capture at once)
            // static staging buffers so the captured graph's pointers stay valid across calls
            if (gx_bytes_ < per_clip * 4 || gout_bytes_ < per_clip * 4) {
                CUDA_CHECK(cudaStreamSynchronize(s));
                cudaFree(gx_); cudaFree(gout_);
                for (auto& g : graphs_) cudaGraphExecDestroy(g.second.exec);
                graphs_.clear();
                CUDA_CHECK(cudaMalloc((void**)&gx_, per_clip * 4));
                CUDA_CHECK(cudaMalloc(&gout_, per_clip * 4));
                gx_bytes_ = gout_bytes_ = per_clip * 4;
            }
            // the caller's stream may be the legacy default stream, which cannot be captured: the graph lives on an
            // engine-owned non-blocking stream, ordered against the caller's stream with events
            if (!gs_) {
                int lo = 0, hi = 0;
                CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));       // hi = most urgent
                CUDA_CHECK(cudaStreamCreateWithPriority(&gs_, cudaStreamNonBlocking, hi));
                CUDA_CHECK(cudaEventCreateWithFlags(&ev_in_, cudaEventDisableTiming));
                CUDA_CHECK(cudaEventCreateWithFlags(&ev_out_, cudaEventDisableTiming));
            }
            ClipGraph& g = graphs_[T];
            if (g.exec && (g.ws != ws || g.out_dtype != out_dtype)) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            CUDA_CHECK(cudaEventRecord(ev_in_, s));
            CUDA_CHECK(cudaStreamWaitEvent(gs_, ev_in_, 0));
            if (!g.exec) {
                cudaGraph_t graph = nullptr;
                const long long l0 = launches_;
                CUDA_CHECK(cudaStreamBeginCapture(gs_, cudaStreamCaptureModeThreadLocal));
                try {
                    begin(ws, ws_bytes, gs_, false);
                    forward_clip(gx_, T, gout_, out_dtype);
                } catch (...) {
                    cudaStreamEndCapture(gs_, &graph);
                    if (graph) cudaGraphDestroy(graph);
                    throw;
                }
                launches_per_clip_[T] = launches_ - l0;   // (counted while enqueueing; the replay below adds them)
                launches_ = l0;
                CUDA_CHECK(cudaStreamEndCapture(gs_, &graph));
                cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
                cudaGraphDestroy(graph);
                CUDA_CHECK(e);
                g.ws = ws; g.out_dtype = out_dtype;
            }
            CUDA_CHECK(cudaMemcpyAsync(gx_, xin, per_clip * 4, cudaMemcpyDeviceToDevice, gs_));
            CUDA_CHECK(cudaGraphLaunch(g.exec, gs_));
            CUDA_CHECK(cudaMemcpyAsync(xout, gout_, per_clip * osz, cudaMemcpyDeviceToDevice, gs_));
            CUDA_CHECK(cudaEventRecord(ev_out_, gs_));
            CUDA_CHECK(cudaStreamWaitEvent(s, ev_out_, 0));
            launches_ += launches_per_clip_[T];
        } else {
            const long long l0 = launches_;
            begin(ws, ws_bytes, s, false);
            forward_clip(xin, T, xout, out_dtype);
            launches_per_clip_[T] = launches_ - l0;
            eager_runs_[T] += 1;
        }
    }
}

// uint8 BGR HWC in / out: the host-side img2tensor / normalize / tensor2img of keep_processor.py:258-260,
// 272-273 folded into the device path -- one conversion kernel in front, the output layout kernel writes the bytes
void Engine::forward_u8(const unsigned char* x_u8_dev, int b, int T, unsigned char* out_u8_dev, void* ws, size_t ws_bytes, cudaStream_t s) {
    KEEP_CHECK(!dry_only_, "keep_forward_u8: engine was created with KEEP_FLAG_PLAN_ONLY (no device)");
    KEEP_CHECK(x_u8_dev && out_u8_dev, "keep_forward_u8: null tensor");
    KEEP_CHECK(b >= 1 && T >= 2 && T <= 100, "keep_forward_u8: need b >= 1 and 2 <= T <= 100 (got b=%d T=%d)", b, T);
    CUDA_CHECK(cudaSetDevice(device_));
    CUDA_CHECK(cudaStreamWaitEvent(s, ev_last_, 0));   // u8_stage_ may still be read by the previous call on another stream
    const size_t need = (size_t)b * T * 3 * 512 * 512 * sizeof(float);
    if (u8_stage_bytes_ < need) {
        CUDA_CHECK(cudaStreamSynchronize(s));
        cudaFree(u8_stage_);
        u8_stage_ = nullptr; u8_stage_bytes_ = 0;
        CUDA_CHECK(cudaMalloc((void**)&u8_stage_, need));
        u8_stage_bytes_ = need;
    }
    u8bgr_to_nchw_norm(x_u8_dev, u8_stage_, b * T, 512, 512, s);
    launches_ += 1;
    forward(u8_stage_, b, T, out_u8_dev, KEEP_OUT_U8_BGR, ws, ws_bytes, s);
}

// =============================================================================================
// profiling (bench.py): CUDA events around every conv/GEMM launch on the launching stream
// =============================================================================================
cudaEvent_t Engine::get_event() {
    if (!ev_pool_.empty()) {
        cudaEvent_t e = ev_pool_.back();
        ev_pool_.pop_back();
        return e;
    }
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreate(&e));
    return e;
}

void Engine::set_profile(bool on) {
    profile_ = on;
    for (auto& p : prof_) { ev_pool_.push_back(p.a); ev_pool_.push_back(p.b); }
    prof_.clear();
}

// out8: [0] launches, [1] ms, [2] GFLOP, [3] GB (algorithmic) for tag 0 (CUDA-core path); [4..7] same for tag 1 (tcgen05)
void Engine::profile_read(double* out8) {
    for (int i = 0; i < 8; ++i) out8[i] = 0.0;
    CUDA_CHECK(cudaDeviceSynchronize());
    for (auto& p : prof_) {
        float ms = 0.0f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, p.a, p.b));
        double* o = out8 + (p.tag ? 4 : 0);
        o[0] += 1.0; o[1] += ms; o[2] += p.flops * 1e-9; o[3] += p.bytes * 1e-9;
    }
}

void Engine::profile_dump(const char* path) {
    CUDA_CHECK(cudaDeviceSynchronize());
    FILE* f = fopen(path, "w");
    KEEP_CHECK(f, "cannot open %s", path);
    fprintf(f, "tag,M,K,N,kh_stride,splitk,bn,ms,gflop\n");
    for (auto& p : prof_) {
        float ms = 0.0f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, p.a, p.b));
        fprintf(f, "%d,%d,%d,%d,%d,%d,%d,%.6f,%.4f\n", p.tag, p.m, p.k, p.n, p.kh, p.splitk, p.bn, ms, p.flops * 1e-9);
    }
    fclose(f);
}

// =============================================================================================
// test hooks
// =============================================================================================
void Engine::force(const std::string& what, const void* host, size_t bytes) {
    CUDA_CHECK(cudaSetDevice(device_));
    Cap& c = forced_[what];
    cudaFree(c.p);
    c.p = nullptr; c.bytes = 0;
    if (!host || bytes == 0) return;
    if (what == "codes") {   // indices gather rows of the 1024-entry codebook
        const int* idx = (const int*)host;
        for (size_t i = 0; i < bytes / sizeof(int); ++i)
            KEEP_CHECK(idx[i] >= 0 && idx[i] < 1024, "forced code index %d at position %zu is outside the codebook", idx[i], i);
    }
    CUDA_CHECK(cudaMalloc(&c.p, bytes));
    CUDA_CHECK(cudaMemcpy(c.p, host, bytes, cudaMemcpyHostToDevice));
    c.bytes = bytes;
}

// sticky non-finite status word: synchronises the device, returns the bits seen since the last clearing read
int Engine::status(bool clear) {
    if (dry_only_ || !status_) return 0;
    CUDA_CHECK(cudaSetDevice(device_));
    CUDA_CHECK(cudaDeviceSynchronize());
    int v = 0;
    CUDA_CHECK(cudaMemcpy(&v, status_, sizeof(int), cudaMemcpyDeviceToHost));
    if (clear && v) CUDA_CHECK(cudaMemset(status_, 0, sizeof(int)));
    return v;
}

size_t Engine::read(const std::string& what, void* host, size_t bytes) {
    CUDA_CHECK(cudaSetDevice(device_));
    auto it = cap_.find(what);
    KEEP_CHECK(it != cap_.end() && it->second.p, "no captured tensor '%s' (enable capture and run forward first)", what.c_str());
    const size_t n = std::min(bytes, it->second.bytes);
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(host, it->second.p, n, cudaMemcpyDeviceToHost));
    return it->second.bytes;
}

}  // namespace keep
