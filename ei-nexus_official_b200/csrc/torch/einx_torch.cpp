// PyTorch operator registration over the C ABI of include/einx.h (the "PyTorch C++/CUDA extension behind a thin
// C-ABI" of the north star): torch.ops.einx.{voxelize, detect, detect_pair, sample, mnn}.
//
// The kernels live in libeinx.so; this shim only does what a PyTorch extension does natively -- tensors in, the
// current CUDA stream of the tensors' device, outputs from the caching allocator, schemas with mutation
// annotations and Meta kernels so that the ops trace (torch.compile / FakeTensor) -- and forwards plain pointers
// and sizes.  One einx_ctx per (device, stream), like the ctypes host layer (a context owns one stream-ordered
// workspace); the Python layer asks this library for its contexts (einx_torch_context) so both bindings share them.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <map>
#include <mutex>
#include <tuple>
#include <utility>

#include "einx.h"

namespace {

std::mutex g_mu;
std::map<std::pair<int, void*>, einx_ctx*> g_ctx;

einx_ctx* context_or_null(int device, void* stream) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto key = std::make_pair(device, stream);
    auto it = g_ctx.find(key);
    if (it != g_ctx.end()) return it->second;
    einx_ctx* ctx = nullptr;
    if (einx_create(device, &ctx) != EINX_OK) return nullptr;  // text: einx_last_error(NULL)
    g_ctx[key] = ctx;
    return ctx;
}

einx_ctx* context_for(int device, void* stream) {
    einx_ctx* ctx = context_or_null(device, stream);
    TORCH_CHECK(ctx != nullptr, "einx_create(", device, ") failed: ", einx_last_error(nullptr),
                " -- there is no CPU or generic-GPU fallback");
    return ctx;
}

struct Call {  // device guard + context + stream of the tensor's device
    c10::cuda::CUDAGuard guard;
    void* stream;
    einx_ctx* ctx;
    explicit Call(const at::Tensor& t) : guard(t.device()) {
        TORCH_CHECK(t.is_cuda(), "einx ops run on CUDA sm_100a only; got a ", t.device(), " tensor (no CPU fallback)");
        stream = (void*)c10::cuda::getCurrentCUDAStream(t.device().index()).stream();
        ctx = context_for(t.device().index(), stream);
    }
    void check(int rc, const char* what) const {
        TORCH_CHECK(rc == EINX_OK, what, " failed (", rc, "): ", einx_last_error(ctx));
    }
};

void want(const at::Tensor& t, at::ScalarType dt, const char* name) {
    TORCH_CHECK(t.scalar_type() == dt && t.is_contiguous(), name, ": expected a contiguous ", dt, " tensor");
}
template <typename T>
T* opt_ptr(const c10::optional<at::Tensor>& t) {
    return (t.has_value() && t->defined()) ? t->data_ptr<T>() : nullptr;
}

// ---- voxelize ---------------------------------------------------------------------------------- //
at::Tensor voxelize_cuda(const at::Tensor& x, const at::Tensor& y, const at::Tensor& t, const at::Tensor& p,
                         const at::Tensor& offsets, int64_t bins, int64_t H, int64_t W, bool normalize) {
    Call c(x);
    want(x, at::kFloat, "x"); want(y, at::kFloat, "y"); want(p, at::kFloat, "p"); want(t, at::kDouble, "t");
    want(offsets, at::kLong, "offsets");
    const int64_t B = offsets.numel() - 1;
    at::Tensor out = at::empty({B, bins, H, W}, x.options());
    c.check(einx_voxelize(c.ctx, x.data_ptr<float>(), y.data_ptr<float>(), t.data_ptr<double>(), p.data_ptr<float>(),
                          offsets.data_ptr<int64_t>(), (int)B, (int)bins, (int)H, (int)W, normalize ? 1 : 0,
                          out.data_ptr<float>(), c.stream), "einx_voxelize");
    return out;
}
at::Tensor voxelize_meta(const at::Tensor& x, const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor& offsets,
                         int64_t bins, int64_t H, int64_t W, bool) {
    return at::empty({offsets.numel() - 1, bins, H, W}, x.options());
}

// ---- detect ------------------------------------------------------------------------------------ //
std::tuple<at::Tensor, at::Tensor, at::Tensor> detect_cuda(at::Tensor score, const c10::optional<at::Tensor>& mask,
                                                           int64_t nms_radius, int64_t border, double prob_thresh,
                                                           int64_t top_k, int64_t kcap, bool want_map) {
    Call c(score);
    want(score, at::kFloat, "score");
    TORCH_CHECK(score.dim() == 4 && score.size(1) == 1 || score.dim() == 3, "score: expected (B, 1, Hp, Wp) or (B, Hp, Wp)");
    const int64_t B = score.size(0), Hp = score.size(-2), Wp = score.size(-1);
    if (mask.has_value() && mask->defined()) want(*mask, at::kByte, "mask");
    at::Tensor kpts = at::empty({B, kcap, 3}, score.options());
    at::Tensor counts = at::empty({B}, score.options().dtype(at::kInt));
    at::Tensor map = want_map ? at::empty_like(score) : at::empty({0}, score.options());
    c.check(einx_detect(c.ctx, score.data_ptr<float>(), opt_ptr<uint8_t>(mask), (int)B, (int)Hp, (int)Wp, (int)nms_radius,
                        (int)border, (float)prob_thresh, (int)top_k, want_map ? map.data_ptr<float>() : nullptr,
                        kpts.data_ptr<float>(), (int)kcap, counts.data_ptr<int32_t>(), c.stream), "einx_detect");
    return {kpts, counts, map};
}
std::tuple<at::Tensor, at::Tensor, at::Tensor> detect_meta(at::Tensor score, const c10::optional<at::Tensor>&, int64_t, int64_t,
                                                           double, int64_t, int64_t kcap, bool want_map) {
    return {at::empty({score.size(0), kcap, 3}, score.options()), at::empty({score.size(0)}, score.options().dtype(at::kInt)),
            want_map ? at::empty_like(score) : at::empty({0}, score.options())};
}

std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor> detect_pair_cuda(
    at::Tensor score0, at::Tensor score1, const c10::optional<at::Tensor>& mask0, const c10::optional<at::Tensor>& mask1,
    int64_t nms_radius, int64_t border, double prob_thresh, int64_t top_k, int64_t kcap) {
    Call c(score0);
    want(score0, at::kFloat, "score0"); want(score1, at::kFloat, "score1");
    TORCH_CHECK(score0.sizes() == score1.sizes() && score0.device() == score1.device(), "detect_pair: the two sides must match in shape and device");
    const int64_t B = score0.size(0), Hp = score0.size(-2), Wp = score0.size(-1);
    at::Tensor k0 = at::empty({B, kcap, 3}, score0.options()), k1 = at::empty({B, kcap, 3}, score0.options());
    at::Tensor c0 = at::empty({B}, score0.options().dtype(at::kInt)), c1 = at::empty({B}, score0.options().dtype(at::kInt));
    c.check(einx_detect_pair(c.ctx, score0.data_ptr<float>(), score1.data_ptr<float>(), opt_ptr<uint8_t>(mask0), opt_ptr<uint8_t>(mask1),
                             (int)B, (int)Hp, (int)Wp, (int)nms_radius, (int)border, (float)prob_thresh, (int)top_k, nullptr, nullptr,
                             k0.data_ptr<float>(), k1.data_ptr<float>(), (int)kcap, c0.data_ptr<int32_t>(), c1.data_ptr<int32_t>(),
                             c.stream), "einx_detect_pair");
    return {k0, c0, k1, c1};
}
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor> detect_pair_meta(at::Tensor score0, at::Tensor, const c10::optional<at::Tensor>&,
                                                                            const c10::optional<at::Tensor>&, int64_t, int64_t, double,
                                                                            int64_t, int64_t kcap) {
    const int64_t B = score0.size(0);
    auto k = [&] { return at::empty({B, kcap, 3}, score0.options()); };
    auto n = [&] { return at::empty({B}, score0.options().dtype(at::kInt)); };
    return {k(), n(), k(), n()};
}

// ---- sample ------------------------------------------------------------------------------------ //
at::Tensor sample_cuda(const at::Tensor& raw, const at::Tensor& kpts, const at::Tensor& counts, int64_t mode, int64_t Hp,
                       int64_t Wp, double scale, bool normalize) {
    Call c(raw);
    TORCH_CHECK(raw.scalar_type() == at::kFloat && raw.dim() == 4, "raw: expected a 4-d float tensor");
    want(kpts, at::kFloat, "kpts"); want(counts, at::kInt, "counts");
    // mode EINX_SAMPLE_GATHER on a channels_last map reads the NHWC memory directly
    int m = (int)mode;
    const int64_t B = raw.size(0), C = raw.size(1), Hd = raw.size(2), Wd = raw.size(3);
    at::Tensor src = raw;
    if (m == EINX_SAMPLE_GATHER && !raw.is_contiguous() && raw.is_contiguous(at::MemoryFormat::ChannelsLast)) m = EINX_SAMPLE_GATHER_NHWC;
    else src = raw.contiguous();
    const int64_t kcap = kpts.size(1);
    at::Tensor desc = at::empty({B, kcap, C}, kpts.options());
    c.check(einx_sample(c.ctx, src.data_ptr<float>(), (int)B, (int)C, (int)Hd, (int)Wd, m, (int)Hp, (int)Wp, kpts.data_ptr<float>(),
                        counts.data_ptr<int32_t>(), (int)kcap, (float)scale, normalize ? 1 : 0, desc.data_ptr<float>(), c.stream),
            "einx_sample");
    return desc;
}
std::tuple<at::Tensor, at::Tensor> sample_split_cuda(const at::Tensor& raw, const at::Tensor& kpts, const at::Tensor& counts, int64_t mode,
                                                     int64_t Hp, int64_t Wp, double scale, bool normalize) {
    Call c(raw);
    TORCH_CHECK(raw.scalar_type() == at::kFloat && raw.dim() == 4, "raw: expected a 4-d float tensor");
    want(kpts, at::kFloat, "kpts"); want(counts, at::kInt, "counts");
    int m = (int)mode;
    const int64_t B = raw.size(0), C = raw.size(1), Hd = raw.size(2), Wd = raw.size(3);
    at::Tensor src = raw;
    if (m == EINX_SAMPLE_GATHER && !raw.is_contiguous() && raw.is_contiguous(at::MemoryFormat::ChannelsLast)) m = EINX_SAMPLE_GATHER_NHWC;
    else src = raw.contiguous();
    const int64_t kcap = kpts.size(1);
    at::Tensor desc = at::empty({B, kcap, C}, kpts.options());
    at::Tensor split = at::empty({2, B, kcap, C}, kpts.options().dtype(at::kHalf));   // [hi | lo] of the FP16X3 matcher mode
    c.check(einx_sample_split(c.ctx, src.data_ptr<float>(), (int)B, (int)C, (int)Hd, (int)Wd, m, (int)Hp, (int)Wp, kpts.data_ptr<float>(),
                              counts.data_ptr<int32_t>(), (int)kcap, (float)scale, normalize ? 1 : 0, desc.data_ptr<float>(),
                              (uint16_t*)split.data_ptr(), c.stream), "einx_sample_split");
    return {desc, split};
}
std::tuple<at::Tensor, at::Tensor> sample_split_meta(const at::Tensor& raw, const at::Tensor& kpts, const at::Tensor&, int64_t, int64_t,
                                                     int64_t, double, bool) {
    return {at::empty({raw.size(0), kpts.size(1), raw.size(1)}, kpts.options()),
            at::empty({2, raw.size(0), kpts.size(1), raw.size(1)}, kpts.options().dtype(at::kHalf))};
}
at::Tensor sample_meta(const at::Tensor& raw, const at::Tensor& kpts, const at::Tensor&, int64_t, int64_t, int64_t, double, bool) {
    return at::empty({raw.size(0), kpts.size(1), raw.size(1)}, kpts.options());
}

// ---- mnn --------------------------------------------------------------------------------------- //
std::vector<at::Tensor> mnn_cuda(const at::Tensor& d0, const at::Tensor& d1, const c10::optional<at::Tensor>& n0,
                                 const c10::optional<at::Tensor>& n1, const c10::optional<at::Tensor>& kpts0,
                                 const c10::optional<at::Tensor>& kpts1, double ratio_thresh, double distance_thresh, bool mutual,
                                 int64_t precision, const c10::optional<at::Tensor>& split0, const c10::optional<at::Tensor>& split1) {
    Call c(d0);
    want(d0, at::kFloat, "d0"); want(d1, at::kFloat, "d1");
    TORCH_CHECK(d0.dim() == 3 && d1.dim() == 3 && d0.size(0) == d1.size(0) && d0.size(2) == d1.size(2), "mnn: (B, N, D) and (B, M, D)");
    const int64_t B = d0.size(0), N = d0.size(1), M = d1.size(1), D = d0.size(2);
    const bool gather = kpts0.has_value() && kpts0->defined();
    auto lo = d0.options().dtype(at::kLong);
    at::Tensor m0 = at::empty({B, N}, lo), m1 = at::empty({B, M}, lo);
    at::Tensor s0 = at::empty({B, N}, d0.options()), s1 = at::empty({B, M}, d0.options());
    at::Tensor mk0, mk1, nm;
    if (gather) {
        TORCH_CHECK(kpts1.has_value() && kpts1->defined(), "mnn: kpts0 given without kpts1");
        want(*kpts0, at::kFloat, "kpts0"); want(*kpts1, at::kFloat, "kpts1");
        TORCH_CHECK(kpts0->size(-1) == 3 && kpts1->size(-1) == 3, "mnn: keypoint rows must be (y, x, prob)");
        mk0 = at::empty({B, N, 3}, d0.options()); mk1 = at::empty({B, N, 3}, d0.options());
        nm = at::empty({B}, d0.options().dtype(at::kInt));
    }
    const uint16_t *sp0 = nullptr, *sp1 = nullptr;
    if (split0.has_value() && split0->defined() && split1.has_value() && split1->defined()) {
        TORCH_CHECK(split0->scalar_type() == at::kHalf && split1->scalar_type() == at::kHalf && split0->is_contiguous() && split1->is_contiguous() &&
                        split0->numel() == 2 * d0.numel() && split1->numel() == 2 * d1.numel(),
                    "mnn: split operands are contiguous (2, B, N|M, D) half tensors (einx::sample_split)");
        sp0 = (const uint16_t*)split0->data_ptr(); sp1 = (const uint16_t*)split1->data_ptr();
    }
    c.check(einx_mnn_split(c.ctx, d0.data_ptr<float>(), d1.data_ptr<float>(), sp0, sp1, opt_ptr<int32_t>(n0), opt_ptr<int32_t>(n1), (int)B, (int)N, (int)M,
                     (int)D, (float)ratio_thresh, (float)distance_thresh, mutual ? 1 : 0, (int)precision, m0.data_ptr<int64_t>(),
                     m1.data_ptr<int64_t>(), s0.data_ptr<float>(), s1.data_ptr<float>(), gather ? kpts0->data_ptr<float>() : nullptr,
                     gather ? kpts1->data_ptr<float>() : nullptr, gather ? mk0.data_ptr<float>() : nullptr,
                     gather ? mk1.data_ptr<float>() : nullptr, gather ? nm.data_ptr<int32_t>() : nullptr, c.stream), "einx_mnn");
    if (gather) return {m0, m1, s0, s1, mk0, mk1, nm};
    return {m0, m1, s0, s1};
}
std::vector<at::Tensor> mnn_meta(const at::Tensor& d0, const at::Tensor& d1, const c10::optional<at::Tensor>&, const c10::optional<at::Tensor>&,
                                 const c10::optional<at::Tensor>& kpts0, const c10::optional<at::Tensor>&, double, double, bool, int64_t,
                                 const c10::optional<at::Tensor>&, const c10::optional<at::Tensor>&) {
    const int64_t B = d0.size(0), N = d0.size(1), M = d1.size(1);
    auto lo = d0.options().dtype(at::kLong);
    std::vector<at::Tensor> out = {at::empty({B, N}, lo), at::empty({B, M}, lo), at::empty({B, N}, d0.options()), at::empty({B, M}, d0.options())};
    if (kpts0.has_value() && kpts0->defined()) {
        out.push_back(at::empty({B, N, 3}, d0.options()));
        out.push_back(at::empty({B, N, 3}, d0.options()));
        out.push_back(at::empty({B}, d0.options().dtype(at::kInt)));
    }
    return out;
}

}  // namespace

// The Python host layer shares these contexts (workspace, launch counter, profiling slots) with the ops above.
extern "C" void* einx_torch_context(int device, void* stream) {
    return context_or_null(device, stream);  // NULL on failure, text via einx_last_error(NULL); never throws
}

// Self-test of the error path (tests/test_abi_and_host.py): a failing C-ABI call must surface as a c10::Error that
// unwinds through this library, with the C-side message attached.  Returns 1 when it does.
extern "C" int einx_torch_error_path_selftest(void) {
    try {
        context_for(1 << 20, nullptr);  // no such device
    } catch (const c10::Error& e) {
        return std::string(e.what()).find("einx_create") != std::string::npos ? 1 : -1;
    } catch (...) {
        return -2;
    }
    return 0;
}

TORCH_LIBRARY(einx, m) {
    m.def("voxelize(Tensor x, Tensor y, Tensor t, Tensor p, Tensor offsets, int bins, int H, int W, bool normalize=True) -> Tensor");
    m.def("detect(Tensor(a!) score, Tensor? mask, int nms_radius, int border, float prob_thresh, int top_k, int kcap, bool want_map=False) -> (Tensor, Tensor, Tensor)");
    m.def("detect_pair(Tensor(a!) score0, Tensor(b!) score1, Tensor? mask0, Tensor? mask1, int nms_radius, int border, float prob_thresh, int top_k, int kcap) -> (Tensor, Tensor, Tensor, Tensor)");
    m.def("sample(Tensor raw, Tensor kpts, Tensor counts, int mode, int Hp, int Wp, float scale, bool normalize=True) -> Tensor");
    m.def("sample_split(Tensor raw, Tensor kpts, Tensor counts, int mode, int Hp, int Wp, float scale, bool normalize=True) -> (Tensor, Tensor)");
    m.def("mnn(Tensor d0, Tensor d1, Tensor? n0, Tensor? n1, Tensor? kpts0, Tensor? kpts1, float ratio_thresh, float distance_thresh, bool mutual, int precision, Tensor? split0=None, Tensor? split1=None) -> Tensor[]");
}
TORCH_LIBRARY_IMPL(einx, CUDA, m) {
    m.impl("voxelize", voxelize_cuda);
    m.impl("detect", detect_cuda);
    m.impl("detect_pair", detect_pair_cuda);
    m.impl("sample", sample_cuda);
    m.impl("sample_split", sample_split_cuda);
    m.impl("mnn", mnn_cuda);
}
TORCH_LIBRARY_IMPL(einx, Meta, m) {
    m.impl("voxelize", voxelize_meta);
    m.impl("detect", detect_meta);
    m.impl("detect_pair", detect_pair_meta);
    m.impl("sample", sample_meta);
    m.impl("sample_split", sample_split_meta);
    m.impl("mnn", mnn_meta);
}
