// wire.cu -- the steps either side of the scoring path: the reference's wire formats and a
// residency table keyed by (segment id, SegmentMeta.State).  Host-side byte handling only; everything numeric stays in
// the kernels behind the entry points this file calls.
//
//  * PQCodebook (vectorsearch.proto:135-142; written by SegmentBuildService.buildCodebookBytes,
//    J/tasks/SegmentBuildService.java:325-338; read by SegmentCaches.decodeCodebook, J/cache/SegmentCaches.java:141-162):
//        int32 m = 1; int32 k = 2; repeated bytes centroids = 3;   // entry s = K * subDim little-endian fp32
//    vs_codebook_encode emits exactly the bytes protobuf-java's toByteArray() produces for that message (fields in
//    number order, zero-valued scalars omitted), vs_codebook_decode accepts any valid encoding of it.
//  * residency: the Java side asks "is segment S resident in state T?" before a query and registers what it uploads;
//    a state change (PENDING -> SEALED: J/tasks/SegmentBuildService.java:100-130; compaction: J/tasks/MaintenanceService.java:
//    388-390 rebuilds through SegmentBuildService.build) or vs_residency_invalidate frees the stale copy.  A byte budget
//    evicts least-recently-used segments.
#include <cstring>
#include <list>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vsgpu.h"
#include "host.h"

namespace vs {
int fail(int code, const char* fmt, ...);
}
using namespace vs;

namespace {

inline size_t varint_len(uint64_t v) {
  size_t n = 1;
  while (v >= 0x80) {
    v >>= 7;
    n++;
  }
  return n;
}
inline uint8_t* put_varint(uint8_t* p, uint64_t v) {
  while (v >= 0x80) {
    *p++ = (uint8_t)(v | 0x80);
    v >>= 7;
  }
  *p++ = (uint8_t)v;
  return p;
}
inline bool get_varint(const uint8_t*& p, const uint8_t* end, uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; shift < 64 && p < end; shift += 7) {
    const uint8_t b = *p++;
    r |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) {
      *v = r;
      return true;
    }
  }
  return false;
}

}  // namespace

extern "C" {

// centroids float[M][K][subDim] -> serialized PQCodebook.  out == NULL (or capacity too small) only reports the length.
int32_t vs_codebook_encode(const float* centroids, int32_t M, int32_t K, int32_t subDim, uint8_t* out, int64_t capacity,
                           int64_t* len_out) {
  if (!centroids || !len_out) return fail(VS_EINVAL, "null pointer");
  if (M <= 0 || K <= 0 || subDim <= 0) return fail(VS_EINVAL, "Invalid PQ params (m,k,dimension)");
  const uint64_t blob = (uint64_t)K * subDim * 4;
  // int32 fields are varint-encoded as sign-extended 64-bit values; both are positive here
  const size_t need = 1 + varint_len((uint64_t)M) + 1 + varint_len((uint64_t)K) + (size_t)M * (1 + varint_len(blob) + blob);
  *len_out = (int64_t)need;
  if (!out) return VS_OK;
  if (capacity < (int64_t)need) return fail(VS_EINVAL, "output holds %lld bytes, the message needs %zu", (long long)capacity, need);
  uint8_t* p = out;
  *p++ = 0x08;  // field 1, varint
  p = put_varint(p, (uint64_t)M);
  *p++ = 0x10;  // field 2, varint
  p = put_varint(p, (uint64_t)K);
  for (int s = 0; s < M; s++) {
    *p++ = 0x1a;  // field 3, length-delimited
    p = put_varint(p, blob);
    // ByteBuffer.order(LITTLE_ENDIAN).putFloat in (ci, di) order = the row-major fp32 bytes on this (little-endian) host
    memcpy(p, centroids + (size_t)s * K * subDim, blob);
    p += blob;
  }
  return VS_OK;
}

// serialized PQCodebook -> M, K, subDim and (if centroids_out != NULL and large enough) float[M][K][subDim].
// Follows SegmentCaches.decodeCodebook: subDim = blob length / (k * 4); additionally rejects what that code would
// mis-read silently (a blob count other than m, blobs of different or non-multiple lengths).
int32_t vs_codebook_decode(const uint8_t* bytes, int64_t len, float* centroids_out, int64_t capacity_floats, int32_t* M_out,
                           int32_t* K_out, int32_t* subDim_out) {
  if (!bytes || len < 0) return fail(VS_EINVAL, "null pointer or negative length");
  const uint8_t *p = bytes, *end = bytes + len;
  int64_t m = 0, k = 0;
  std::vector<std::pair<const uint8_t*, uint64_t>> blobs;
  while (p < end) {
    uint64_t tag, v;
    if (!get_varint(p, end, &tag)) return fail(VS_EINVAL, "truncated PQCodebook message");
    const uint32_t field = (uint32_t)(tag >> 3), type = (uint32_t)(tag & 7);
    if (type == 0) {
      if (!get_varint(p, end, &v)) return fail(VS_EINVAL, "truncated PQCodebook message");
      if (field == 1) m = (int32_t)v;
      if (field == 2) k = (int32_t)v;
    } else if (type == 2) {
      if (!get_varint(p, end, &v) || (uint64_t)(end - p) < v) return fail(VS_EINVAL, "truncated PQCodebook message");
      if (field == 3) blobs.emplace_back(p, v);
      p += v;
    } else if (type == 1) {
      if (end - p < 8) return fail(VS_EINVAL, "truncated PQCodebook message");
      p += 8;
    } else if (type == 5) {
      if (end - p < 4) return fail(VS_EINVAL, "truncated PQCodebook message");
      p += 4;
    } else {
      return fail(VS_EINVAL, "unsupported wire type %u in PQCodebook message", type);
    }
  }
  if (m <= 0 || k <= 0) return fail(VS_EINVAL, "PQCodebook has m = %lld, k = %lld", (long long)m, (long long)k);
  if ((int64_t)blobs.size() != m) return fail(VS_EINVAL, "PQCodebook has m = %lld but %zu centroid blobs", (long long)m, blobs.size());
  const uint64_t blob = blobs[0].second;
  if (blob == 0 || blob % ((uint64_t)k * 4) != 0) return fail(VS_EINVAL, "centroid blob of %llu bytes is not k * subDim * 4", (unsigned long long)blob);
  for (auto& b : blobs)
    if (b.second != blob) return fail(VS_EINVAL, "centroid blobs differ in length");
  const int64_t sub = (int64_t)(blob / ((uint64_t)k * 4));
  if (M_out) *M_out = (int32_t)m;
  if (K_out) *K_out = (int32_t)k;
  if (subDim_out) *subDim_out = (int32_t)sub;
  if (!centroids_out) return VS_OK;
  if (capacity_floats < m * k * sub) return fail(VS_EINVAL, "output holds %lld floats, the codebook has %lld", (long long)capacity_floats, (long long)(m * k * sub));
  for (int64_t s = 0; s < m; s++) memcpy(centroids_out + (size_t)s * k * sub, blobs[s].first, blob);
  return VS_OK;
}

// Sealing from stored bytes: the PQCodebook message and (nullable) the codes -> vs_segment_attach_pq
int32_t vs_segment_attach_pq_codebook(uint64_t h, const uint8_t* codebook, int64_t len, const uint8_t* codes) {
  int32_t M = 0, K = 0, sub = 0;
  int r = vs_codebook_decode(codebook, len, nullptr, 0, &M, &K, &sub);
  if (r != VS_OK) return r;
  int32_t d = 0;
  r = vs_segment_info(h, nullptr, &d, nullptr, nullptr, nullptr);
  if (r != VS_OK) return r;
  if ((int64_t)M * sub != d) return fail(VS_EINVAL, "codebook is %d x %d floats per vector, the segment has dimension %d", M, sub, d);
  std::vector<float> cent((size_t)M * K * sub);
  r = vs_codebook_decode(codebook, len, cent.data(), (int64_t)cent.size(), nullptr, nullptr, nullptr);
  if (r != VS_OK) return r;
  return vs_segment_attach_pq(h, cent.data(), M, K, codes);
}

// ---- residency table ---------------------------------------------------------------------------------------------
namespace {
struct Resident {
  uint64_t handle;
  int32_t state;
  int64_t bytes;
  std::list<int64_t>::iterator lru;
};
std::mutex g_res_mu;
std::unordered_map<int64_t, Resident> g_res;
std::list<int64_t> g_lru;  // front = most recently used
int64_t g_res_bytes = 0, g_res_budget = 0;  // budget 0 = unlimited

int64_t handle_bytes(uint64_t h) {
  int64_t n = 0;
  int32_t d = 0, M = 0, K = 0;
  if (vs_segment_info(h, &n, &d, &M, &K, nullptr) != VS_OK) return -1;
  // rows, codes, codebook, and what the first queries add (coefficients, the fp16 operand copy: api.cu)
  return n * d * 4 + n * M + (int64_t)K * d * 4 + nomination_aux_bytes(n, d);
}
}  // namespace

// Registers `handle` as the resident copy of segment seg_id in `state` (SegmentMeta.State: 0 ACTIVE, 1 PENDING, 2 SEALED,
// 3 COMPACTING, 4 WRITING).  A previous copy of the same segment is freed; least-recently-used segments are freed until
// the table fits its budget again (the new entry itself is never evicted).  The table owns the handles it holds.
int32_t vs_residency_put(int64_t seg_id, int32_t state, uint64_t handle) {
  if (state < 0 || state > 4) return fail(VS_EINVAL, "unknown SegmentMeta.State %d", state);
  const int64_t bytes = handle_bytes(handle);
  if (bytes < 0) return fail(VS_EHANDLE, "unknown segment handle");
  std::vector<uint64_t> to_free;
  {
    std::lock_guard<std::mutex> g(g_res_mu);
    auto it = g_res.find(seg_id);
    if (it != g_res.end()) {
      if (it->second.handle != handle) to_free.push_back(it->second.handle);
      g_res_bytes -= it->second.bytes;
      g_lru.erase(it->second.lru);
      g_res.erase(it);
    }
    g_lru.push_front(seg_id);
    g_res[seg_id] = Resident{handle, state, bytes, g_lru.begin()};
    g_res_bytes += bytes;
    while (g_res_budget > 0 && g_res_bytes > g_res_budget && g_lru.size() > 1) {
      const int64_t victim = g_lru.back();
      auto vt = g_res.find(victim);
      to_free.push_back(vt->second.handle);
      g_res_bytes -= vt->second.bytes;
      g_lru.pop_back();
      g_res.erase(vt);
    }
  }
  for (uint64_t h : to_free) vs_segment_free(h);
  return VS_OK;
}

// The resident copy of seg_id, if there is one IN THAT STATE: a copy made while the segment was PENDING is stale once
// the segment is SEALED (it has no codes) -- VS_ESTATE tells the caller to upload / attach and put again.
int32_t vs_residency_get(int64_t seg_id, int32_t state, uint64_t* handle_out) {
  if (!handle_out) return fail(VS_EINVAL, "null pointer");
  std::lock_guard<std::mutex> g(g_res_mu);
  auto it = g_res.find(seg_id);
  if (it == g_res.end()) return fail(VS_EHANDLE, "segment %lld is not resident", (long long)seg_id);
  if (it->second.state != state) {
    *handle_out = it->second.handle;  // still usable for an upgrade in place (attach PQ, then put with the new state)
    return fail(VS_ESTATE, "segment %lld is resident in state %d, not %d", (long long)seg_id, it->second.state, state);
  }
  g_lru.erase(it->second.lru);
  g_lru.push_front(seg_id);
  it->second.lru = g_lru.begin();
  *handle_out = it->second.handle;
  return VS_OK;
}

// Compaction rebuild, deletion of a segment, vacuum: drop the resident copy (MaintenanceService.java:388-390)
int32_t vs_residency_invalidate(int64_t seg_id) {
  uint64_t h = 0;
  {
    std::lock_guard<std::mutex> g(g_res_mu);
    auto it = g_res.find(seg_id);
    if (it == g_res.end()) return VS_OK;  // nothing resident: nothing to do
    h = it->second.handle;
    g_res_bytes -= it->second.bytes;
    g_lru.erase(it->second.lru);
    g_res.erase(it);
  }
  return vs_segment_free(h);
}

int32_t vs_residency_set_budget(int64_t bytes) {
  if (bytes < 0) return fail(VS_EINVAL, "budget must be >= 0 (0 = unlimited)");
  std::lock_guard<std::mutex> g(g_res_mu);
  g_res_budget = bytes;
  return VS_OK;
}

int32_t vs_residency_stats(int64_t* segments_out, int64_t* bytes_out) {
  std::lock_guard<std::mutex> g(g_res_mu);
  if (segments_out) *segments_out = (int64_t)g_res.size();
  if (bytes_out) *bytes_out = g_res_bytes;
  return VS_OK;
}

}  // extern "C"

namespace vs {
// vs_shutdown: the handles are about to be freed wholesale
void residency_clear() {
  std::lock_guard<std::mutex> g(g_res_mu);
  g_res.clear();
  g_lru.clear();
  g_res_bytes = 0;
}
}  // namespace vs
