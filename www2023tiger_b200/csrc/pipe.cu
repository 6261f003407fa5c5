// Host-side batch pipeline (no device code): per staging slot one captured finder graph and one captured model
// graph, uploads + finder on a copy-in stream, model graphs in batch order on the caller's stream, downloads on a
// copy-out stream.  This is what www2023tiger_b200/engine.py:StreamRunner drives; doing the seven CUDA calls of a
// step here instead of through a Python framework keeps the host ~4x ahead of a 75 us GPU step.
// Reference counterpart: the training loop's per-batch `.to(device)` / `.item()` traffic
// (train_self_supervised.py:140-175), which is synchronous there.
#include <vector>

#include "common.cuh"

struct TigerPipe {
  int n_slots;
  cudaStream_t s_in, s_out;
  std::vector<cudaGraphExec_t> finder, model, tail;
  std::vector<cudaEvent_t> ev_in, ev_done, ev_out;
  std::vector<char> used, has_out;
  int last_slot;
};

extern "C" void* tiger_pipe_create(int n_slots) {
  if (n_slots <= 0 || n_slots > 1024) return nullptr;
  TigerPipe* p = new TigerPipe();
  p->n_slots = n_slots;
  if (cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking) != cudaSuccess) {
    delete p;
    return nullptr;
  }
  p->finder.assign(n_slots, nullptr);
  p->model.assign(n_slots, nullptr);
  p->tail.assign(n_slots, nullptr);
  p->has_out.assign(n_slots, 0);
  p->last_slot = -1;
  p->ev_in.resize(n_slots);
  p->ev_done.resize(n_slots);
  p->ev_out.resize(n_slots);
  p->used.assign(n_slots, 0);
  for (int i = 0; i < n_slots; ++i) {
    cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&p->ev_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming);
  }
  return p;
}

extern "C" void tiger_pipe_destroy(void* pipe) {
  TigerPipe* p = reinterpret_cast<TigerPipe*>(pipe);
  if (p == nullptr) return;
  for (int i = 0; i < p->n_slots; ++i) {
    if (p->finder[i] != nullptr) cudaGraphExecDestroy(p->finder[i]);
    if (p->model[i] != nullptr) cudaGraphExecDestroy(p->model[i]);
    if (p->tail[i] != nullptr) cudaGraphExecDestroy(p->tail[i]);
    cudaEventDestroy(p->ev_in[i]);
    cudaEventDestroy(p->ev_done[i]);
    cudaEventDestroy(p->ev_out[i]);
  }
  cudaStreamDestroy(p->s_in);
  cudaStreamDestroy(p->s_out);
  delete p;
}

// Everything launched on `stream` (and on streams that join it through events) between begin and end becomes the
// slot's finder (kind 0), model (kind 1) or tail (kind 2, optional) graph.  The tail graph holds kernels that only
// produce the batch's results and change no state (the link scorer): it is replayed on the copy-out stream in front
// of the download, i.e. beside the next batch's model kernels.
extern "C" int tiger_pipe_capture_begin(void* stream) {
  return cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeRelaxed) == cudaSuccess ? TIGER_OK : TIGER_ECUDA;
}

extern "C" int tiger_pipe_capture_end(void* pipe, void* stream, int slot, int kind) {
  TigerPipe* p = reinterpret_cast<TigerPipe*>(pipe);
  cudaGraph_t graph = nullptr;
  if (cudaStreamEndCapture(as_stream(stream), &graph) != cudaSuccess || graph == nullptr) return TIGER_ECUDA;
  if (p == nullptr || slot < 0 || slot >= p->n_slots || kind < 0 || kind > 2) {
    cudaGraphDestroy(graph);
    return TIGER_EINVAL;
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t err = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (err != cudaSuccess) return TIGER_ECUDA;
  cudaGraphExec_t& dst = kind == 0 ? p->finder[slot] : (kind == 1 ? p->model[slot] : p->tail[slot]);
  if (dst != nullptr) cudaGraphExecDestroy(dst);
  dst = exec;
  return TIGER_OK;
}

// One batch: src (pinned host or device memory, in_bytes) -> d_in on the copy-in stream, the slot's finder graph
// behind it; the slot's model graph on `main_stream` (batch order = call order); d_out -> h_out (out_bytes, may be
// NULL / 0) on the copy-out stream.  The slot's previous batch must have been waited for (host path) - its
// device-side consumption is also awaited on the copy-in stream before d_in is overwritten.
extern "C" int tiger_pipe_submit(void* pipe, int slot, const void* src, void* d_in, int64_t in_bytes, const void* d_out,
                                 void* h_out, int64_t out_bytes, void* main_stream) {
  TigerPipe* p = reinterpret_cast<TigerPipe*>(pipe);
  if (p == nullptr || slot < 0 || slot >= p->n_slots || p->finder[slot] == nullptr || p->model[slot] == nullptr ||
      src == nullptr || d_in == nullptr || in_bytes <= 0)
    return TIGER_EINVAL;
  cudaStream_t main = as_stream(main_stream);
  bool ok = true;
  // the slot's buffers are free once its previous batch has been consumed: by the model graph, or - when a tail
  // graph / download followed - by those
  if (p->used[slot])
    ok = ok && cudaStreamWaitEvent(p->s_in, p->has_out[slot] ? p->ev_out[slot] : p->ev_done[slot], 0) == cudaSuccess;
  ok = ok && cudaMemcpyAsync(d_in, src, (size_t)in_bytes, cudaMemcpyDefault, p->s_in) == cudaSuccess;
  ok = ok && cudaGraphLaunch(p->finder[slot], p->s_in) == cudaSuccess;
  ok = ok && cudaEventRecord(p->ev_in[slot], p->s_in) == cudaSuccess;
  ok = ok && cudaStreamWaitEvent(main, p->ev_in[slot], 0) == cudaSuccess;
  ok = ok && cudaGraphLaunch(p->model[slot], main) == cudaSuccess;
  ok = ok && cudaEventRecord(p->ev_done[slot], main) == cudaSuccess;
  p->used[slot] = 1;
  p->last_slot = slot;
  const bool download = h_out != nullptr && out_bytes > 0;
  p->has_out[slot] = (p->tail[slot] != nullptr || download) ? 1 : 0;
  if (p->has_out[slot]) {
    ok = ok && cudaStreamWaitEvent(p->s_out, p->ev_done[slot], 0) == cudaSuccess;
    if (p->tail[slot] != nullptr) ok = ok && cudaGraphLaunch(p->tail[slot], p->s_out) == cudaSuccess;
    if (download) ok = ok && cudaMemcpyAsync(h_out, d_out, (size_t)out_bytes, cudaMemcpyDefault, p->s_out) == cudaSuccess;
    ok = ok && cudaEventRecord(p->ev_out[slot], p->s_out) == cudaSuccess;
  }
  return ok ? TIGER_OK : TIGER_ECUDA;
}

// Blocks until the slot's results are in host memory (after a submit with h_out) / its model graph has finished.
extern "C" int tiger_pipe_wait(void* pipe, int slot, int host_results) {
  TigerPipe* p = reinterpret_cast<TigerPipe*>(pipe);
  if (p == nullptr || slot < 0 || slot >= p->n_slots) return TIGER_EINVAL;
  const bool out = host_results && p->has_out[slot];
  return cudaEventSynchronize(out ? p->ev_out[slot] : p->ev_done[slot]) == cudaSuccess ? TIGER_OK : TIGER_ECUDA;
}

// Makes `stream` wait for everything the most recent submit put on the copy-out stream (tail graph, download).
extern "C" int tiger_pipe_join(void* pipe, void* stream) {
  TigerPipe* p = reinterpret_cast<TigerPipe*>(pipe);
  if (p == nullptr) return TIGER_EINVAL;
  if (p->last_slot < 0 || !p->has_out[p->last_slot]) return TIGER_OK;
  return cudaStreamWaitEvent(as_stream(stream), p->ev_out[p->last_slot], 0) == cudaSuccess ? TIGER_OK : TIGER_ECUDA;
}
