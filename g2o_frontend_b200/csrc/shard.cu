// shard.cu -- pair-wise sharding of batched alignment over the GPUs of one process.
//
// The caller that owns this loop in the reference is the loop closer: PwnCloser::processPartition walks the candidate
// pairs of a partition one by one, builds (or fetches from its cache) the clouds of both frames and matches them
// (pwn_tracker2/pwn_closer.cpp:83-182 -> PwnMatcherBase::makeCloud / matchClouds, pwn_matcher_base.cpp:46-196).  Pairs are
// independent, so the pair list is cut into contiguous blocks, one per device; every device builds the clouds of the
// frames ITS pairs reference from the raw images (a frame needed on two devices is prepared on both: 50 MB of local
// traffic instead of 25 MB over NVLink) and aligns its block; the only "collective" is the result records landing in the
// caller's array.  No data-path exchange, hence no NCCL here: one process, host memory is shared.  (The one-process-per-
// GPU deployment gathers the same 256-byte records with one all-gather, g2o_frontend_b200/sharding.py.)
//
// One worker thread per device per call; each worker owns a nicp_context (one context per host thread / GPU, like the
// rest of the library) and a cache of device clouds that persists across calls.
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "nicp_internal.cuh"

struct nicp_shard_pool {
  struct Worker {
    int device;
    nicp_context *ctx;
    std::vector<nicp_cloud *> clouds;  // cache, capacity = cloudCapacity each
    int cloudCapacity;
    int rc;
    std::string error;
  };
  std::vector<Worker> workers;
};

using namespace nicp;

extern "C" {

int nicp_shard_pool_create(const int *devices, int n_devices, nicp_shard_pool **out) {
  if (!out) return NICP_ERR_INVALID;
  *out = nullptr;
  if (!devices || n_devices <= 0) {
    set_error("a shard pool needs at least one device");
    return NICP_ERR_INVALID;
  }
  nicp_shard_pool *pool = new nicp_shard_pool();
  pool->workers.resize(n_devices);
  for (int i = 0; i < n_devices; i++) {
    nicp_shard_pool::Worker &w = pool->workers[i];
    w.device = devices[i];
    w.ctx = nullptr;
    w.cloudCapacity = 0;
    w.rc = NICP_OK;
    int rc = nicp_create(devices[i], &w.ctx);  // fails loudly without a usable GPU: there is no CPU fallback
    if (rc != NICP_OK) {
      for (int j = 0; j < i; j++) nicp_destroy(pool->workers[j].ctx);
      delete pool;
      return rc;
    }
  }
  *out = pool;
  return NICP_OK;
}

void nicp_shard_pool_destroy(nicp_shard_pool *pool) {
  if (!pool) return;
  for (nicp_shard_pool::Worker &w : pool->workers) {
    for (nicp_cloud *c : w.clouds) nicp_cloud_destroy(c);
    nicp_destroy(w.ctx);
  }
  delete pool;
}

int nicp_shard_pool_size(const nicp_shard_pool *pool) { return pool ? (int)pool->workers.size() : 0; }

int nicp_align_frames_sharded(nicp_shard_pool *pool, int n_frames, const uint16_t *const *raws, int raw_rows, int raw_cols,
                              float depth_scale, int step, float max_depth_cov, const nicp_projector *proj,
                              const nicp_stats_params *sp, const float sensor_offset[16], int n_pairs,
                              const int *reference_frame, const int *current_frame, const float *initial_guesses,
                              const nicp_align_params *ap, float frame_inlier_depth_threshold, nicp_align_result *results) {
  if (!pool || n_frames < 0 || n_pairs < 0 || !proj || !sp || !ap || raw_rows <= 0 || raw_cols <= 0) return NICP_ERR_INVALID;
  if (n_pairs == 0) return NICP_OK;
  if (!raws || !reference_frame || !current_frame || !results) return NICP_ERR_INVALID;
  for (int i = 0; i < n_pairs; i++)
    if (reference_frame[i] < 0 || reference_frame[i] >= n_frames || current_frame[i] < 0 || current_frame[i] >= n_frames) {
      set_error("pair %d references frame %d / %d outside [0, %d)", i, reference_frame[i], current_frame[i], n_frames);
      return NICP_ERR_INVALID;
    }
  for (int f = 0; f < n_frames; f++)
    if (!raws[f]) {
      set_error("frame %d is null", f);
      return NICP_ERR_INVALID;
    }
  const int G = (int)pool->workers.size();
  const int px = proj->rows * proj->cols;

  auto work = [&](int g) {
    nicp_shard_pool::Worker &w = pool->workers[g];
    w.rc = NICP_OK;
    w.error.clear();
    // static block partition of the pair list: device g takes pairs [g n / G, (g + 1) n / G)
    const long long lo = (long long)g * n_pairs / G, hi = (long long)(g + 1) * n_pairs / G;
    const int m = (int)(hi - lo);
    if (m <= 0) return;
    auto fail = [&](int rc) {
      w.rc = rc;
      w.error = nicp_last_error();  // this worker thread's message
    };
    // the frames this block references, in order of first use -> local cloud slots
    std::map<int, int> slotOf;
    std::vector<int> frames;
    for (long long i = lo; i < hi; i++)
      for (int f : {current_frame[i], reference_frame[i]})
        if (slotOf.find(f) == slotOf.end()) {
          slotOf[f] = (int)frames.size();
          frames.push_back(f);
        }
    if (w.cloudCapacity < px) {  // a different image size: the cache starts over
      for (nicp_cloud *c : w.clouds) nicp_cloud_destroy(c);
      w.clouds.clear();
      w.cloudCapacity = px;
    }
    while (w.clouds.size() < frames.size()) {
      nicp_cloud *c = nullptr;
      int rc = nicp_cloud_create(w.ctx, w.cloudCapacity, &c);
      if (rc) return fail(rc);
      w.clouds.push_back(c);
    }
    std::vector<const uint16_t *> r(frames.size());
    for (size_t k = 0; k < frames.size(); k++) r[k] = raws[frames[k]];
    int rc = nicp_raw_depth_to_cloud_batch(w.ctx, (int)frames.size(), r.data(), raw_rows, raw_cols, depth_scale, step,
                                           max_depth_cov, proj, sp, sensor_offset, 0, w.clouds.data());
    if (rc) return fail(rc);
    std::vector<const nicp_cloud *> refs(m), curs(m);
    for (int i = 0; i < m; i++) {
      refs[i] = w.clouds[slotOf[reference_frame[lo + i]]];
      curs[i] = w.clouds[slotOf[current_frame[lo + i]]];
    }
    rc = nicp_align_batch(w.ctx, m, refs.data(), curs.data(), proj, ap, sensor_offset, sensor_offset,
                          initial_guesses ? initial_guesses + 16 * lo : nullptr, frame_inlier_depth_threshold, results + lo);
    if (rc) return fail(rc);
  };

  if (G == 1) {
    work(0);
  } else {
    std::vector<std::thread> threads;
    threads.reserve(G);
    for (int g = 0; g < G; g++) threads.emplace_back(work, g);
    for (std::thread &t : threads) t.join();
  }
  for (int g = 0; g < G; g++)
    if (pool->workers[g].rc != NICP_OK) {
      set_error("device %d (shard %d of %d): %s", pool->workers[g].device, g, G, pool->workers[g].error.c_str());
      return pool->workers[g].rc;
    }
  return NICP_OK;
}

}  // extern "C"
