// batch.cu — chips_detect over a batch of images without Python ("plus batched
// variants with a leading B dim"; §8(e): a time series, images are independent).
//
// The reference has no batching: one `Chips` object handles one `RegisterAIA`
// (/root/reference/chips/chips.py:110-150) and a caller loops over dates.  chips_detect_batch is
// that loop for a non-Python host: B host (or device) images go through `n_slots` in-flight
// detections, each slot with its own CUDA stream, device image, workspace and - per geometry - an
// instantiated CUDA graph of the detection's ~200 launches (enqueueing them one by one costs the
// host more than the GPU needs per image).  Per image only the rectangle the kernels read is
// uploaded (chips_input_rect); back come the 1.6 kB record and, on request, the level-encoded
// result map (map_t = keep <= t).  The optional summed probability map (plots.py:305-322 over the
// batch) is accumulated on one stream in image order, so it is reproducible.
#include <new>
#include <vector>
#include "common.cuh"

namespace {

struct Slot {
  cudaStream_t stream = nullptr;
  void* image = nullptr;          // device image, H x W x elem
  void* ws = nullptr;             // workspace of plan_cap.bytes
  void* rec_pinned = nullptr;     // pinned staging of the record
  chips_plan plan{};
  int cx = 0, cy = 0, r = 0;
  bool has_graph = false;
  cudaGraphExec_t exec = nullptr;
  double key[4] = {0, 0, 0, 0};   // ht_peak_ratio, h_thresh_in, porb, area of the captured graph
  std::vector<double> offsets;
  std::vector<int32_t> spans;     // chips_input_spans of the captured geometry
  cudaEvent_t detected = nullptr; // detection of the slot's current image has finished
  cudaEvent_t summed = nullptr;   // ... and its probability map is in the sum
  int pending = -1;               // batch index whose results are still on the device
};

struct Batch {
  int H, W, k, h_bins, T, elem, masked, n_slots;
  chips_plan plan_cap{};          // largest workspace any geometry of this shape needs
  std::vector<Slot> slots;
  cudaStream_t sum_stream = nullptr;
  float* prob_sum = nullptr;      // device, H x W
};

void destroy(Batch* b) {
  if (!b) return;
  for (Slot& s : b->slots) {
    if (s.exec) cudaGraphExecDestroy(s.exec);
    if (s.detected) cudaEventDestroy(s.detected);
    if (s.summed) cudaEventDestroy(s.summed);
    if (s.rec_pinned) cudaFreeHost(s.rec_pinned);
    if (s.ws) cudaFree(s.ws);
    if (s.image) cudaFree(s.image);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  if (b->prob_sum) cudaFree(b->prob_sum);
  if (b->sum_stream) cudaStreamDestroy(b->sum_stream);
  delete b;
}

bool same_call(const Slot& s, int cx, int cy, int r, const double* key, const double* offs, int T) {
  if (!s.has_graph || s.cx != cx || s.cy != cy || s.r != r) return false;
  for (int i = 0; i < 4; ++i)
    if (!(s.key[i] == key[i]) && !(s.key[i] != s.key[i] && key[i] != key[i])) return false;   // NaN == NaN here
  for (int i = 0; i < T; ++i)
    if (s.offsets[i] != offs[i]) return false;
  return true;
}

}  // namespace

extern "C" int chips_batch_create(void** handle, int H, int W, int k, int h_bins, int T, int elem,
                                  int masked, int n_slots) {
  if (!handle || n_slots < 1 || n_slots > 64) return CHIPS_E_ARG;
  *handle = nullptr;
  Batch* b = new (std::nothrow) Batch();
  if (!b) return CHIPS_E_WS;
  b->H = H; b->W = W; b->k = k; b->h_bins = h_bins; b->T = T; b->elem = elem;
  b->masked = masked != 0; b->n_slots = n_slots;
  // a disk that covers the whole image has the largest ROI, hence the largest workspace
  const int rcap = masked ? H + W : -1;
  int rc = chips_plan_init(&b->plan_cap, H, W, k, h_bins, T, elem, W / 2, H / 2, rcap);
  if (rc) { delete b; return rc; }
  b->slots.resize((size_t)n_slots);
  bool ok = cudaStreamCreateWithFlags(&b->sum_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&b->prob_sum, (size_t)H * W * sizeof(float)) == cudaSuccess;
  for (Slot& s : b->slots) {
    if (!ok) break;
    ok = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess &&
         cudaMalloc(&s.image, (size_t)H * W * elem) == cudaSuccess &&
         cudaMalloc(&s.ws, b->plan_cap.bytes) == cudaSuccess &&
         cudaMallocHost(&s.rec_pinned, sizeof(chips_result)) == cudaSuccess &&
         cudaEventCreateWithFlags(&s.detected, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&s.summed, cudaEventDisableTiming) == cudaSuccess;
    s.offsets.assign((size_t)(T > 0 ? T : 1), 0.0);
  }
  if (!ok) {
    destroy(b);
    cudaGetLastError();
    return CHIPS_E_CUDA;
  }
  *handle = b;
  return CHIPS_OK;
}

extern "C" int chips_batch_destroy(void* handle) {
  destroy(static_cast<Batch*>(handle));
  return CHIPS_OK;
}

extern "C" int chips_detect_batch(void* handle, int B, const void* const* images, int images_on_device,
                                  const int32_t* cx, const int32_t* cy, const int32_t* r,
                                  double ht_peak_ratio, double h_thresh_in, const double* offsets,
                                  double porb_threshold, double area_threshold, chips_result* results,
                                  uint8_t* const* keep_or_null, float* prob_sum_or_null,
                                  double prob_lower_lim) {
  Batch* b = static_cast<Batch*>(handle);
  if (!b || B < 0 || (B > 0 && (!images || !cx || !cy || !r || !offsets || !results))) return CHIPS_E_ARG;
  const size_t npix = (size_t)b->H * b->W, pitch = (size_t)b->W * b->elem;
  const double key[4] = {ht_peak_ratio, h_thresh_in, porb_threshold, area_threshold};
  if (prob_sum_or_null) CHIPS_CUDA(cudaMemsetAsync(b->prob_sum, 0, npix * sizeof(float), b->sum_stream));

  // results of the image a slot still holds -> host (called before the slot is reused, and at the end)
  auto collect = [&](Slot& s) -> int {
    if (s.pending < 0) return CHIPS_OK;
    const int i = s.pending;
    unsigned char* w = static_cast<unsigned char*>(s.ws);
    CHIPS_CUDA(cudaMemcpyAsync(s.rec_pinned, w + s.plan.off_result, sizeof(chips_result),
                               cudaMemcpyDeviceToHost, s.stream));
    if (keep_or_null && keep_or_null[i])
      CHIPS_CUDA(cudaMemcpyAsync(keep_or_null[i], w + s.plan.off_keep, npix, cudaMemcpyDeviceToHost, s.stream));
    CHIPS_CUDA(cudaStreamSynchronize(s.stream));
    results[i] = *static_cast<const chips_result*>(s.rec_pinned);
    if (prob_sum_or_null) CHIPS_CUDA(cudaEventSynchronize(s.summed));   // the workspace is free again
    s.pending = -1;
    return CHIPS_OK;
  };

  for (int i = 0; i < B; ++i) {
    Slot& s = b->slots[(size_t)(i % b->n_slots)];
    int rc = collect(s);
    if (rc) return rc;
    if (!images[i] || (b->masked ? r[i] < 0 : r[i] >= 0)) return CHIPS_E_ARG;
    if (!same_call(s, cx[i], cy[i], r[i], key, offsets, b->T)) {
      rc = chips_plan_init(&s.plan, b->H, b->W, b->k, b->h_bins, b->T, b->elem, cx[i], cy[i], r[i]);
      if (rc) return rc;
      if (s.plan.bytes > b->plan_cap.bytes) return CHIPS_E_WS;
      if (s.exec) {
        cudaGraphExecDestroy(s.exec);
        s.exec = nullptr;
      }
      s.has_graph = false;
      cudaGraph_t graph = nullptr;
      CHIPS_CUDA(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
      rc = chips_detect(s.image, &s.plan, cx[i], cy[i], r[i], ht_peak_ratio, h_thresh_in, offsets,
                        porb_threshold, area_threshold, nullptr, s.ws, s.stream);
      const cudaError_t ce = cudaStreamEndCapture(s.stream, &graph);
      if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return rc ? rc : CHIPS_E_CUDA;
      }
      const cudaError_t ci = cudaGraphInstantiate(&s.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ci != cudaSuccess) return CHIPS_E_CUDA;
      s.spans.assign(4 * 4096, 0);
      const int ns = chips_input_spans(&s.plan, cx[i], cy[i], r[i], s.spans.data(), 4096);
      if (ns < 1) return ns < 0 ? ns : CHIPS_E_ARG;
      s.spans.resize((size_t)4 * ns);
      s.cx = cx[i]; s.cy = cy[i]; s.r = r[i];
      for (int j = 0; j < 4; ++j) s.key[j] = key[j];
      for (int j = 0; j < b->T; ++j) s.offsets[(size_t)j] = offsets[j];
      s.has_graph = true;
    }
    // upload the row intervals the kernels read (chips_input_spans)
    for (size_t q = 0; q + 3 < s.spans.size(); q += 4) {
      const int y0 = s.spans[q], y1 = s.spans[q + 1], x0 = s.spans[q + 2], x1 = s.spans[q + 3];
      const size_t off = (size_t)y0 * pitch + (size_t)x0 * b->elem;
      CHIPS_CUDA(cudaMemcpy2DAsync(static_cast<char*>(s.image) + off, pitch,
                                   static_cast<const char*>(images[i]) + off, pitch,
                                   (size_t)(x1 - x0) * b->elem, (size_t)(y1 - y0),
                                   images_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s.stream));
    }
    CHIPS_CUDA(cudaGraphLaunch(s.exec, s.stream));
    chips::count_launch(0);
    CHIPS_CUDA(cudaEventRecord(s.detected, s.stream));
    if (prob_sum_or_null) {   // summed in image order on one stream
      CHIPS_CUDA(cudaStreamWaitEvent(b->sum_stream, s.detected, 0));
      rc = chips_prob_stack(&s.plan, prob_lower_lim, b->prob_sum, 4, 1, s.ws, b->sum_stream);
      if (rc) return rc;
      CHIPS_CUDA(cudaEventRecord(s.summed, b->sum_stream));
    }
    s.pending = i;
  }
  for (Slot& s : b->slots) {
    int rc = collect(s);
    if (rc) return rc;
  }
  if (prob_sum_or_null) {
    CHIPS_CUDA(cudaMemcpyAsync(prob_sum_or_null, b->prob_sum, npix * sizeof(float), cudaMemcpyDeviceToHost,
                               b->sum_stream));
    CHIPS_CUDA(cudaStreamSynchronize(b->sum_stream));
  }
  return CHIPS_OK;
}
