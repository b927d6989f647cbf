// CUDA-graph replay of a training step.  The step's C calls (nerfca_train_step, nerfca_adam_step / nerfca_allreduce_adam_step,
// nerfca_gather_batch, nerfca_jitter_depth) are bracketed by nerfca_graph_begin / nerfca_graph_end_launch on a non-default
// stream: the launches are captured into a graph instead of being executed, the executable graph of the previous step is
// updated in place with the new kernel parameters (pointers and per-step scalars may change freely; only a changed launch
// sequence forces a re-instantiation) and launched once.  One driver submission per step instead of one per kernel, and the
// kernels run back to back without host-side launch gaps.
#include "common.cuh"

// A small rotation of executable graphs, each with the event of its last launch: an executable graph is only updated once its
// previous launch has finished (the host waits on that event), which also bounds the number of steps the host can run ahead of the
// GPU to NERFCA_GRAPH_DEPTH.  (Updating ONE executable graph while thousands of its launches were still queued crashed inside the
// driver in a 4 000-step soak run.)
constexpr int NERFCA_GRAPH_DEPTH = 4;
struct nerfca_graph {
  cudaGraphExec_t exec[NERFCA_GRAPH_DEPTH] = {};
  cudaEvent_t done[NERFCA_GRAPH_DEPTH] = {};
  bool capturing = false;
  long long launches = 0, updates = 0, instantiations = 0;
};

using namespace nerfca;

extern "C" int nerfca_graph_create(nerfca_graph** out) {
  NERFCA_REQUIRE(out != nullptr, NERFCA_E_ARG, "null pointer");
  *out = new nerfca_graph();
  return NERFCA_OK;
}

extern "C" int nerfca_graph_begin(nerfca_graph* g, void* stream) {
  NERFCA_REQUIRE(g != nullptr, NERFCA_E_ARG, "null graph");
  NERFCA_REQUIRE(stream != nullptr, NERFCA_E_ARG, "graph capture needs a non-default stream");
  NERFCA_REQUIRE(!g->capturing, NERFCA_E_ARG, "nerfca_graph_begin called twice");
  NERFCA_CUDA_OK(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
  g->capturing = true;
  return NERFCA_OK;
}

extern "C" int nerfca_graph_end_launch(nerfca_graph* g, void* stream) {
  NERFCA_REQUIRE(g != nullptr && g->capturing, NERFCA_E_ARG, "nerfca_graph_end_launch without nerfca_graph_begin");
  g->capturing = false;
  cudaGraph_t graph = nullptr;
  NERFCA_CUDA_OK(cudaStreamEndCapture((cudaStream_t)stream, &graph));
  const int k = (int)(g->launches % NERFCA_GRAPH_DEPTH);
  if (g->done[k]) {
    cudaError_t e = cudaEventSynchronize(g->done[k]);            // the launch that last used this executable graph has finished
    if (e != cudaSuccess) {
      cudaGraphDestroy(graph);
      set_error(std::string("nerfca_graph_end_launch: a previous step failed -> ") + cudaGetErrorString(e));
      return NERFCA_E_CUDA;
    }
  } else {
    NERFCA_CUDA_OK(cudaEventCreateWithFlags(&g->done[k], cudaEventDisableTiming));
  }
  bool ok = false;
  if (g->exec[k]) {
    cudaGraphExecUpdateResultInfo info;
    if (cudaGraphExecUpdate(g->exec[k], graph, &info) == cudaSuccess) {
      ok = true;
      ++g->updates;
    } else {
      cudaGetLastError();                      // a changed launch sequence: build a new executable graph
      cudaGraphExecDestroy(g->exec[k]);
      g->exec[k] = nullptr;
    }
  }
  if (!ok) {
    cudaError_t e = cudaGraphInstantiate(&g->exec[k], graph, 0);
    if (e != cudaSuccess) {
      cudaGraphDestroy(graph);
      set_error(std::string("nerfca_graph_end_launch: cudaGraphInstantiate -> ") + cudaGetErrorString(e));
      return NERFCA_E_CUDA;
    }
    ++g->instantiations;
  }
  cudaGraphDestroy(graph);
  NERFCA_CUDA_OK(cudaGraphLaunch(g->exec[k], (cudaStream_t)stream));
  NERFCA_CUDA_OK(cudaEventRecord(g->done[k], (cudaStream_t)stream));
  ++g->launches;
  return NERFCA_OK;
}

// Abandon a capture after a failed call inside the bracket (the stream leaves capture mode, nothing is launched).
extern "C" int nerfca_graph_abort(nerfca_graph* g, void* stream) {
  NERFCA_REQUIRE(g != nullptr, NERFCA_E_ARG, "null graph");
  if (g->capturing) {
    cudaGraph_t graph = nullptr;
    cudaStreamEndCapture((cudaStream_t)stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    g->capturing = false;
  }
  return NERFCA_OK;
}

extern "C" int nerfca_graph_stats(const nerfca_graph* g, int64_t* launches, int64_t* updates, int64_t* instantiations) {
  NERFCA_REQUIRE(g != nullptr, NERFCA_E_ARG, "null graph");
  if (launches) *launches = g->launches;
  if (updates) *updates = g->updates;
  if (instantiations) *instantiations = g->instantiations;
  return NERFCA_OK;
}

extern "C" int nerfca_graph_destroy(nerfca_graph* g) {
  if (!g) return NERFCA_OK;
  for (int k = 0; k < NERFCA_GRAPH_DEPTH; ++k) {
    if (g->exec[k]) cudaGraphExecDestroy(g->exec[k]);
    if (g->done[k]) cudaEventDestroy(g->done[k]);
  }
  delete g;
  return NERFCA_OK;
}
