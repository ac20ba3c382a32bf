// native_reducer.cpp — the Reducer's per-parameter hook loop in C++ (a torch extension; the default hook loop
// when it is built, B200DDP_NATIVE=0 selects the Python loop of reducer.py).
//
// The reference's DDP runs this loop natively (c10d::Reducer: autograd_hook -> mark_variable_ready ->
// mark_bucket_ready -> all_reduce_bucket -> finalize_backward, torch:include/torch/csrc/distributed/c10d/
// reducer.hpp:74,280,282,524,286); reducer.py restates it in Python, this file restates the same state
// machine without the interpreter on the autograd thread: a post hook on every parameter's AccumulateGrad
// node counts the bucket's pending gradients down; a complete bucket is launched — through the SAME C ABI
// (include/b200ddp.h, function pointers handed over by the Python host) — as soon as every earlier bucket has
// been; an engine callback ends the iteration.  No torch type crosses into libb200ddp.so.
//
// Bucket layout, arena offsets and algorithms are decided by the Python host (bucketing.py / P2PEngine.plan)
// and installed with configure(); the one-time rebuild stays in Python (it needs the host-side order sync).
#include <torch/extension.h>

#include <torch/csrc/autograd/engine.h>
#include <torch/csrc/autograd/utils/lambda_post_hook.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

using fn_stream_wait = int (*)(void*, void*, void*);
using fn_flatten = int (*)(void*, const void* const*, const uint64_t*, const uint64_t*, int, int, uint64_t, float, void*);
using fn_unflatten = int (*)(void*, void* const*, const uint64_t*, const uint64_t*, int, int, uint64_t, void*);
using fn_allreduce = int (*)(void*, uint64_t, uint64_t, uint64_t, int, int, int, float, void*);
using fn_fused = int (*)(void*, void* const*, const uint64_t*, const uint64_t*, int, int, uint64_t, uint64_t, float, void*);
using fn_unflatten_sgd = int (*)(void*, void* const*, void* const*, const uint64_t*, const uint64_t*, int, int, uint64_t,
                                 float, void*);
using fn_fused_sgd = int (*)(void*, void* const*, void* const*, const uint64_t*, const uint64_t*, int, int, uint64_t,
                             uint64_t, float, float, void*);
using fn_poll = int (*)(void*);
using fn_last_error = const char* (*)(void*);

struct Bucket {
  std::vector<int64_t> params;
  std::vector<uint64_t> offs, lens;
  int dtype = 0;
  uint64_t nbytes = 0, arena_off = 0, out_off = 0;
  int algo = 0;
  bool fused = false;
  // gradient_as_bucket_view (reducer.hpp:238,328-334): views of the bucket in the arena, one per parameter; when
  // present the bucket is not unflattened, the parameters' .grad are re-pointed at these views instead
  std::vector<at::Tensor> views;
  // per iteration
  int64_t pending = 0;
  bool ready = false;
};

class NativeReducer {
 public:
  NativeReducer(uintptr_t ctx, const std::map<std::string, uintptr_t>& fns, int64_t world, uintptr_t comm_stream,
                int64_t max_ctas, std::vector<at::Tensor> params)
      : ctx_(reinterpret_cast<void*>(ctx)),
        world_(world),
        comm_stream_(reinterpret_cast<void*>(comm_stream)),
        max_ctas_(static_cast<int>(max_ctas)),
        params_(std::move(params)) {
    auto need = [&](const char* name) {
      auto it = fns.find(name);
      if (it == fns.end() || it->second == 0) throw std::runtime_error(std::string("missing C-ABI entry point ") + name);
      return it->second;
    };
    stream_wait_ = reinterpret_cast<fn_stream_wait>(need("b200ddp_stream_wait"));
    flatten_ = reinterpret_cast<fn_flatten>(need("b200ddp_flatten_scale"));
    unflatten_ = reinterpret_cast<fn_unflatten>(need("b200ddp_unflatten"));
    allreduce_ = reinterpret_cast<fn_allreduce>(need("b200ddp_allreduce"));
    fused_ = reinterpret_cast<fn_fused>(need("b200ddp_fused_bucket"));
    {  // optional: the optimizer-epilogue entry points (absent from the recording tables of the CPU tests)
      auto u = fns.find("b200ddp_unflatten_sgd");
      auto f = fns.find("b200ddp_fused_bucket_sgd");
      if (u != fns.end() && f != fns.end()) {
        unflatten_sgd_ = reinterpret_cast<fn_unflatten_sgd>(u->second);
        fused_sgd_ = reinterpret_cast<fn_fused_sgd>(f->second);
      }
    }
    poll_ = reinterpret_cast<fn_poll>(need("b200ddp_poll_error"));
    last_error_ = reinterpret_cast<fn_last_error>(need("b200ddp_last_error"));
    scale_ = 1.0f / static_cast<float>(world_);
    param_bucket_.assign(params_.size(), -1);
    seen_.assign(params_.size(), 0);
    for (size_t i = 0; i < params_.size(); ++i)
      if (!params_[i].requires_grad()) throw std::runtime_error("every reduced parameter must require a gradient");
    install_hooks();
  }

  // (Re-)create the AccumulateGrad nodes and hang the post hooks on them.  An AccumulateGrad node remembers the
  // stream that was current when it was CREATED and the engine runs it there; torch's own Reducer keeps its nodes
  // for good, which is why stock DDP must be constructed under the stream later forwards use (and why a DDP built
  // on the default stream cannot be captured into a CUDA graph).  We instead re-create the nodes whenever the
  // forward's stream changes (DDP.forward -> Reducer.bind_stream): eager steps, a capture on torch's capture
  // stream, eager steps again — each with nodes of its own stream.  Only legal between iterations.
  void rehook() {
    std::lock_guard<std::mutex> lk(mu_);
    if (armed_) throw std::runtime_error("cannot re-create the gradient accumulators while a backward is in flight");
    remove_hooks();  // drops our references: nodes no graph holds any more are destroyed here
    install_hooks();
  }

  ~NativeReducer() { remove_hooks(); }

  // the engine re-created its context (arena regrow): later launches go to the new one
  void rebind(uintptr_t ctx) {
    std::lock_guard<std::mutex> lk(mu_);
    if (armed_) throw std::runtime_error("cannot rebind the context while a backward is in flight");
    ctx_ = reinterpret_cast<void*>(ctx);
  }

  // buckets: list of (param_indices, offsets_el, lengths_el, dtype_code, nbytes, arena_off, out_off, algo, fused,
  //                   views)   — views: [] or one arena view per parameter (gradient_as_bucket_view)
  void configure(const std::vector<std::tuple<std::vector<int64_t>, std::vector<uint64_t>, std::vector<uint64_t>, int64_t,
                                              uint64_t, uint64_t, uint64_t, int64_t, bool, std::vector<at::Tensor>>>& buckets) {
    std::lock_guard<std::mutex> lk(mu_);
    if (armed_) throw std::runtime_error("cannot re-configure buckets while a backward is in flight");
    buckets_.clear();
    std::fill(param_bucket_.begin(), param_bucket_.end(), -1);
    for (size_t b = 0; b < buckets.size(); ++b) {
      Bucket bk;
      std::tie(bk.params, bk.offs, bk.lens, bk.dtype, bk.nbytes, bk.arena_off, bk.out_off, bk.algo, bk.fused) =
          std::make_tuple(std::get<0>(buckets[b]), std::get<1>(buckets[b]), std::get<2>(buckets[b]),
                          static_cast<int>(std::get<3>(buckets[b])), std::get<4>(buckets[b]), std::get<5>(buckets[b]),
                          std::get<6>(buckets[b]), static_cast<int>(std::get<7>(buckets[b])), std::get<8>(buckets[b]));
      bk.views = std::get<9>(buckets[b]);
      if (bk.params.size() != bk.offs.size() || bk.params.size() != bk.lens.size())
        throw std::runtime_error("bucket tables differ in length");
      if (!bk.views.empty() && (bk.views.size() != bk.params.size() || bk.fused))
        throw std::runtime_error("a bucket served as views needs one view per parameter and the three-kernel path");
      for (int64_t p : bk.params) {
        if (p < 0 || p >= static_cast<int64_t>(params_.size()) || param_bucket_[p] != -1)
          throw std::runtime_error("parameter index out of range or assigned to two buckets");
        param_bucket_[p] = static_cast<int64_t>(b);
      }
      buckets_.push_back(std::move(bk));
    }
    for (int64_t pb : param_bucket_)
      if (pb < 0) throw std::runtime_error("a parameter is not assigned to any bucket");
  }

  // Reducer::prepare_for_backward (called after the forward, DDP._post_forward, distributed.py:1648)
  void prepare() {
    std::lock_guard<std::mutex> lk(mu_);
    if (unfinished_)
      throw std::runtime_error(
          "Expected to have finished reduction in the prior iteration before starting a new one. Some parameters did "
          "not receive a gradient; find_unused_parameters is not supported.");
    for (auto& b : buckets_) {
      b.pending = static_cast<int64_t>(b.params.size());
      b.ready = false;
    }
    std::fill(seen_.begin(), seen_.end(), 0);
    next_bucket_ = 0;
    callback_queued_ = false;
    armed_ = true;
  }

  void disarm() {
    std::lock_guard<std::mutex> lk(mu_);
    armed_ = false;
  }

  // optimizer epilogue for the coming backward: lr > 0 makes the kernels that write the averaged gradient back also
  // apply p <- p - lr g (DistributedDataParallel.fuse_optimizer decides per iteration); 0 = off
  void set_sgd(double lr) {
    std::lock_guard<std::mutex> lk(mu_);
    if (lr != 0.0 && (!unflatten_sgd_ || !fused_sgd_)) throw std::runtime_error("the optimizer-epilogue entry points are missing");
    sgd_lr_ = static_cast<float>(lr);
  }

  // an armed iteration that will never finish (aborted graph capture): forget it
  void abort() {
    std::lock_guard<std::mutex> lk(mu_);
    armed_ = false;
    unfinished_ = false;
    callback_queued_ = false;
    keepalive_.clear();
  }

  std::vector<int64_t> take_ready_order() {
    std::lock_guard<std::mutex> lk(mu_);
    record_order_ = false;
    return std::move(ready_order_);
  }
  int64_t iterations() const { return iterations_; }
  int64_t launches() const { return launches_; }
  std::vector<int64_t> launched_buckets() const { return launched_; }

  void remove_hooks() {
    for (auto& h : hooks_)
      if (h.second) h.second->del_post_hook(h.first);
    hooks_.clear();
  }

 private:
  void install_hooks() {
    for (size_t i = 0; i < params_.size(); ++i) {
      auto acc = torch::autograd::impl::grad_accumulator(params_[i]);
      const uintptr_t key = acc->add_post_hook(std::make_unique<torch::autograd::utils::LambdaPostHook>(
          [this, i](const torch::autograd::variable_list& outputs, const torch::autograd::variable_list&) {
            this->autograd_hook(i);
            return outputs;
          }));
      hooks_.emplace_back(key, acc);  // keeps the accumulator (and so the hook) alive, like Reducer::hooks_
    }
  }

  [[noreturn]] void fail(const char* what, int rc) {
    const char* msg = last_error_ ? last_error_(ctx_) : "";
    throw std::runtime_error(std::string("b200ddp ") + what + " failed (code " + std::to_string(rc) + "): " + (msg ? msg : ""));
  }

  static void* current_stream(const at::Tensor& t) {
    if (!t.is_cuda()) return nullptr;  // CPU tensors only occur in the host-logic tests
    return c10::cuda::getCurrentCUDAStream(t.device().index()).stream();
  }

  void autograd_hook(size_t index) {
    std::lock_guard<std::mutex> lk(mu_);
    if (!armed_) return;
    if (!callback_queued_) {
      callback_queued_ = true;
      unfinished_ = true;
      torch::autograd::Engine::get_default_engine().queue_callback([this] { this->finalize(); });
    }
    if (seen_[index])
      throw std::runtime_error("a parameter was marked ready twice in one backward (re-entrant backward is not supported)");
    seen_[index] = 1;
    if (record_order_) ready_order_.push_back(static_cast<int64_t>(index));
    Bucket& mine = buckets_[param_bucket_[index]];
    if (--mine.pending == 0) {
      mine.ready = true;
      while (next_bucket_ < buckets_.size() && buckets_[next_bucket_].ready) launch(next_bucket_++);
    }
  }

  // one bucket: order the side stream after the autograd stream, then flatten -> allreduce -> unflatten
  // (or the single fused launch) through the C ABI
  void launch(size_t b) {
    Bucket& bk = buckets_[b];
    const size_t n = bk.params.size();
    ptrs_.resize(n);
    std::vector<at::Tensor> dense;  // contiguous stand-ins for exotic gradient layouts
    for (size_t k = 0; k < n; ++k) {
      const at::Tensor& g = params_[bk.params[k]].grad();
      if (!g.defined()) throw std::runtime_error("a ready parameter has no gradient");
      if (g.is_contiguous()) {
        ptrs_[k] = g.data_ptr();
      } else {
        dense.push_back(g.contiguous());
        ptrs_[k] = dense.back().data_ptr();
      }
    }
    void* cur = current_stream(params_[bk.params[0]]);
    int rc = stream_wait_(ctx_, comm_stream_, cur);
    if (rc) fail("stream_wait", rc);
    const bool sgd = sgd_lr_ != 0.0f && bk.views.empty();
    if (sgd) {
      pptrs_.resize(n);
      for (size_t k = 0; k < n; ++k) pptrs_[k] = params_[bk.params[k]].data_ptr();
    }
    if (bk.fused) {
      rc = sgd ? fused_sgd_(ctx_, ptrs_.data(), pptrs_.data(), bk.offs.data(), bk.lens.data(), static_cast<int>(n),
                            bk.dtype, bk.arena_off, bk.nbytes, scale_, sgd_lr_, comm_stream_)
               : fused_(ctx_, ptrs_.data(), bk.offs.data(), bk.lens.data(), static_cast<int>(n), bk.dtype, bk.arena_off,
                        bk.nbytes, scale_, comm_stream_);
      if (rc) fail("fused_bucket", rc);
      ++launches_;
    } else {
      rc = flatten_(ctx_, ptrs_.data(), bk.offs.data(), bk.lens.data(), static_cast<int>(n), bk.dtype, bk.arena_off,
                    scale_, comm_stream_);
      if (rc) fail("flatten_scale", rc);
      if (world_ > 1) {
        rc = allreduce_(ctx_, bk.arena_off, bk.out_off, bk.nbytes, bk.dtype, bk.algo, max_ctas_, 1.0f, comm_stream_);
        if (rc) fail("allreduce", rc);
      }
      if (bk.views.empty()) {
        rc = sgd ? unflatten_sgd_(ctx_, ptrs_.data(), pptrs_.data(), bk.offs.data(), bk.lens.data(), static_cast<int>(n),
                                  bk.dtype, bk.out_off, sgd_lr_, comm_stream_)
                 : unflatten_(ctx_, ptrs_.data(), bk.offs.data(), bk.lens.data(), static_cast<int>(n), bk.dtype,
                              bk.out_off, comm_stream_);
        if (rc) fail("unflatten", rc);
        launches_ += world_ > 1 ? 3 : 2;
      } else {
        // no copy back: .grad becomes the view of the reduced bucket; the tensors the flatten kernel still reads
        // on the side stream stay alive until finalize
        for (size_t k = 0; k < n; ++k) {
          at::Tensor& g = params_[bk.params[k]].mutable_grad();
          if (g.data_ptr() != bk.views[k].data_ptr()) {
            keepalive_.push_back(g);
            g = bk.views[k];
          }
        }
        launches_ += world_ > 1 ? 2 : 1;
      }
    }
    if (!dense.empty() && !bk.views.empty()) {
      for (auto& t : dense) keepalive_.push_back(std::move(t));  // only read by the flatten kernel
    } else if (!dense.empty()) {
      // copy the reduced dense stand-ins back on the side stream and keep them alive until finalize
      size_t d = 0;
      if (params_[bk.params[0]].is_cuda()) {
        c10::cuda::CUDAStreamGuard guard(c10::cuda::getStreamFromExternal(static_cast<cudaStream_t>(comm_stream_),
                                                                          params_[bk.params[0]].device().index()));
        for (size_t k = 0; k < n; ++k) {
          at::Tensor g = params_[bk.params[k]].grad();
          if (!g.is_contiguous()) g.copy_(dense[d++]);
        }
      } else {
        for (size_t k = 0; k < n; ++k) {
          at::Tensor g = params_[bk.params[k]].grad();
          if (!g.is_contiguous()) g.copy_(dense[d++]);
        }
      }
      for (auto& t : dense) keepalive_.push_back(std::move(t));
    }
    launched_.push_back(static_cast<int64_t>(b));
  }

  // Reducer::finalize_backward: the autograd stream waits for the side stream, errors surface here
  void finalize() {
    std::lock_guard<std::mutex> lk(mu_);
    if (!armed_) return;
    armed_ = false;
    void* cur = current_stream(params_[0]);
    int rc = stream_wait_(ctx_, cur, comm_stream_);
    keepalive_.clear();
    if (rc) fail("stream_wait", rc);
    if (next_bucket_ != buckets_.size())
      throw std::runtime_error("backward finished but some parameters never produced a gradient; "
                               "find_unused_parameters is not supported");
    rc = poll_(ctx_);
    if (rc) fail("kernel", rc);
    unfinished_ = false;
    ++iterations_;
  }

  void* ctx_;
  int64_t world_;
  void* comm_stream_;
  int max_ctas_;
  float scale_ = 1.0f;
  std::vector<at::Tensor> params_;
  std::vector<std::pair<uintptr_t, std::shared_ptr<torch::autograd::Node>>> hooks_;
  std::vector<Bucket> buckets_;
  std::vector<int64_t> param_bucket_;
  std::vector<char> seen_;
  std::vector<void*> ptrs_, pptrs_;
  float sgd_lr_ = 0.0f;
  fn_unflatten_sgd unflatten_sgd_ = nullptr;
  fn_fused_sgd fused_sgd_ = nullptr;
  std::vector<at::Tensor> keepalive_;
  std::vector<int64_t> ready_order_, launched_;
  size_t next_bucket_ = 0;
  bool armed_ = false, callback_queued_ = false, unfinished_ = false, record_order_ = true;
  int64_t iterations_ = 0, launches_ = 0;
  std::mutex mu_;
  fn_stream_wait stream_wait_ = nullptr;
  fn_flatten flatten_ = nullptr;
  fn_unflatten unflatten_ = nullptr;
  fn_allreduce allreduce_ = nullptr;
  fn_fused fused_ = nullptr;
  fn_poll poll_ = nullptr;
  fn_last_error last_error_ = nullptr;
};

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "b200ddp native Reducer hook loop (see csrc/native_reducer.cpp)";
  py::class_<NativeReducer>(m, "NativeReducer")
      .def(py::init<uintptr_t, const std::map<std::string, uintptr_t>&, int64_t, uintptr_t, int64_t, std::vector<at::Tensor>>(),
           py::arg("ctx"), py::arg("fns"), py::arg("world"), py::arg("comm_stream"), py::arg("max_ctas"), py::arg("params"))
      .def("configure", &NativeReducer::configure)
      .def("rebind", &NativeReducer::rebind)
      .def("rehook", &NativeReducer::rehook)
      .def("prepare", &NativeReducer::prepare, py::call_guard<py::gil_scoped_release>())
      .def("disarm", &NativeReducer::disarm)
      .def("abort", &NativeReducer::abort)
      .def("set_sgd", &NativeReducer::set_sgd)
      .def("take_ready_order", &NativeReducer::take_ready_order)
      .def("iterations", &NativeReducer::iterations)
      .def("launches", &NativeReducer::launches)
      .def("launched_buckets", &NativeReducer::launched_buckets)
      .def("remove_hooks", &NativeReducer::remove_hooks);
}
