// C ABI entry points (include/femgpu.h): host-side bookkeeping that mirrors the reference's
// FEM<V> container — node numbering, duplicate checks, error texts — plus staging of the
// struct-of-arrays element data into HBM. All numerics run on the device (prep.cu, symbolic.cu,
// numeric.cu); there is no CPU fallback.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

using namespace femgpu;

namespace {

Status g_create_status;  // errors before a handle exists

// Rust's `{:?}` for f64 (used inside the reference's error messages): shortest round-trip digits,
// plain decimal with at least one fractional digit for 1e-4 <= |x| < 1e16, scientific otherwise.
std::string rust_debug_f64(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
  if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
  char buf[64];
  int prec = 1;
  for (; prec <= 17; ++prec) {
    snprintf(buf, sizeof buf, "%.*e", prec - 1, v);
    if (strtod(buf, nullptr) == v) break;
  }
  // buf = d.ddddde[+-]XX
  std::string s(buf);
  size_t epos = s.find('e');
  std::string mant = s.substr(0, epos);
  int exp10 = atoi(s.c_str() + epos + 1);
  bool neg = mant[0] == '-';
  if (neg) mant = mant.substr(1);
  std::string digits;
  for (char c : mant)
    if (c != '.') digits += c;
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  std::string out;
  double a = std::fabs(v);
  if (a >= 1e-4 && a < 1e16) {
    if (exp10 >= 0) {
      std::string ip = digits.substr(0, std::min<size_t>(digits.size(), size_t(exp10) + 1));
      while (ip.size() < size_t(exp10) + 1) ip += '0';
      std::string fp = digits.size() > size_t(exp10) + 1 ? digits.substr(size_t(exp10) + 1) : "0";
      out = ip + "." + fp;
    } else {
      out = "0." + std::string(size_t(-exp10 - 1), '0') + digits;
    }
  } else {
    out = digits.substr(0, 1);
    if (digits.size() > 1) out += "." + digits.substr(1);
    out += "e" + std::to_string(exp10);
  }
  return neg ? "-" + out : out;
}

inline uint64_t bits_of(double v) {
  if (v == 0.0) v = 0.0;  // -0.0 == 0.0 in the reference's `==`
  uint64_t u;
  std::memcpy(&u, &v, 8);
  return u;
}

const char* family_name(int f) { return f == FEMGPU_TRUSS ? "Truss" : (f == FEMGPU_BEAM ? "Beam" : "Plate"); }

// Texts of the element-level checks (structs/truss.rs:29-42, structs/beam.rs:37-61,
// structs/plate.rs:37-56). The reference prints young_modulus inside the Poisson (beam, plate) and
// Thickness (plate) messages; that quirk is kept.
std::string element_error_text(const Handle* h, int family, size_t i, int code) {
  const FamilyHost& f = h->fh[family];
  auto P = [&](int k) { return rust_debug_f64(f.props[k][i]); };
  uint32_t number = f.number[i];
  switch (code) {
    case FEMGPU_E_YOUNG_MODULUS: return "Young's modulus " + P(0) + " is less or equal to zero!";
    case FEMGPU_E_POISSON_RATIO: return "Poisson's ratio " + P(0) + " is less or equal to zero!";
    case FEMGPU_E_AREA: return "Area " + P(family == FEMGPU_TRUSS ? 1 : 2) + " is less or equal to zero!";
    case FEMGPU_E_AREA2: return "Area2 " + P(2) + " is less or equal to zero!";
    case FEMGPU_E_I11: return "I11 " + P(3) + " is less or equal to zero!";
    case FEMGPU_E_I22: return "I22 " + P(4) + " is less or equal to zero!";
    case FEMGPU_E_IT: return "It " + P(6) + " is less or equal to zero!";
    case FEMGPU_E_SHEAR_FACTOR:
      return "Shear factor " + P(family == FEMGPU_BEAM ? 7 : 3) + " is less or equal to zero!";
    case FEMGPU_E_PARALLEL_LOCAL_AXIS:
      return "Local axis 1 direction [" + P(8) + ", " + P(9) + ", " + P(10) + "] parallel to element " +
             std::to_string(number) + "!";
    case FEMGPU_E_THICKNESS: return "Thickness " + P(0) + " is less or equal to zero!";
    case FEMGPU_E_NODES_ON_LINE: return "Some nodes of " + std::to_string(number) + " element lie on the line!";
    case FEMGPU_E_NODES_NOT_ON_PLANE:
      return "Not all nodes of element " + std::to_string(number) + " lie on the plane!";
    case FEMGPU_E_NOT_CONVEX: return "Element " + std::to_string(number) + " non-convex!";
    default: return "element error " + std::to_string(code);
  }
}

// host-side half of the property checks: the sign tests, in the reference's order
int property_check(int family, const double* p /* props of one element */) {
  switch (family) {
    case FEMGPU_TRUSS:
      if (p[0] <= 0.0) return FEMGPU_E_YOUNG_MODULUS;
      if (p[1] <= 0.0) return FEMGPU_E_AREA;
      if (!std::isnan(p[2]) && p[2] <= 0.0) return FEMGPU_E_AREA2;
      return 0;
    case FEMGPU_BEAM:
      if (p[0] <= 0.0) return FEMGPU_E_YOUNG_MODULUS;
      if (p[1] <= 0.0) return FEMGPU_E_POISSON_RATIO;
      if (p[2] <= 0.0) return FEMGPU_E_AREA;
      if (p[3] <= 0.0) return FEMGPU_E_I11;
      if (p[4] <= 0.0) return FEMGPU_E_I22;
      if (p[6] <= 0.0) return FEMGPU_E_IT;
      if (p[7] <= 0.0) return FEMGPU_E_SHEAR_FACTOR;
      return 0;
    default:
      if (p[0] <= 0.0) return FEMGPU_E_YOUNG_MODULUS;
      if (p[1] <= 0.0) return FEMGPU_E_POISSON_RATIO;
      if (p[2] <= 0.0) return FEMGPU_E_THICKNESS;
      if (p[3] <= 0.0) return FEMGPU_E_SHEAR_FACTOR;
      return 0;
  }
}

// the model changed: everything derived from the old one is stale — the pattern, the assembled values, the
// separated matrix and whatever was solved / recovered from it
void invalidate(Handle* h) {
  h->symbolic_valid = false;
  h->values_valid = false;
  h->nz_valid = false;
  h->sep.valid = false;
  h->sep.sky_valid = false;
  sol_invalidate(h);
}

inline void sort4(uint32_t v[4]) {
  auto cs = [&](int a, int b) {
    if (v[a] > v[b]) std::swap(v[a], v[b]);
  };
  cs(0, 1); cs(2, 3); cs(0, 2); cs(1, 3); cs(1, 2);
}

// hash of an element's node-index set (order-insensitive): Truss/Beam::is_nodes_numbers_same accept
// either orientation (structs/truss.rs:257-260), Plate::is_nodes_numbers_same any permutation
// (structs/plate.rs:1114-1118)
inline uint64_t nodeset_hash(int family, const uint32_t nd[4]) {
  if (family != FEMGPU_PLATE) return mix64((uint64_t(std::min(nd[0], nd[1])) << 32) | std::max(nd[0], nd[1]));
  uint32_t v[4] = {nd[0], nd[1], nd[2], nd[3]};
  sort4(v);
  return mix64(((uint64_t(v[0]) << 32) | v[1]) * 0x9E3779B97F4A7C15ull ^ ((uint64_t(v[2]) << 32) | v[3]));
}

inline bool nodeset_equal(int family, const uint32_t a[4], const uint32_t b[4]) {
  if (family != FEMGPU_PLATE)
    return (a[0] == b[0] && a[1] == b[1]) || (a[0] == b[1] && a[1] == b[0]);
  uint32_t x[4] = {a[0], a[1], a[2], a[3]}, y[4] = {b[0], b[1], b[2], b[3]};
  sort4(x);
  sort4(y);
  return x[0] == y[0] && x[1] == y[1] && x[2] == y[2] && x[3] == y[3];
}

inline uint64_t xyz_hash(double x, double y, double z) {
  uint64_t h = mix64(bits_of(x));
  h = mix64(h ^ (bits_of(y) + 0x9E3779B97F4A7C15ull));
  h = mix64(h ^ (bits_of(z) + 0xC2B2AE3D27D4EB4Full));
  return h;
}

// ---- host staging registered with CUDA ---------------------------------------------------------------------------
// A re-used instance (FEM::reset, then the next model) fills the same host vectors again: from its first reset on
// the handle registers them with CUDA (cudaHostRegister, once per allocation), so an upload is a DMA straight out of
// the staging — no bounce through the pinned chunks, no host thread busy — and can be queued right when an add_*
// call returns, under the host checks of the next one. Rules that keep this safe: a vector is unregistered (after a
// stream synchronisation) before it may reallocate (make_room) and before the handle goes; whoever rewrites staging
// that an upload may still be reading (reset, rollback) synchronises the stream first.
bool pinned_is(Handle* h, const void* base, size_t bytes) {
  for (const auto& r : h->pinned)
    if (r.base == base) return r.bytes == bytes;
  return false;
}

void pin_drop(Handle* h, const void* base) {
  for (size_t k = 0; k < h->pinned.size(); ++k)
    if (h->pinned[k].base == base) {
      cudaSetDevice(h->device);
      cudaStreamSynchronize(h->stream);
      if (cudaHostUnregister(const_cast<void*>(base)) != cudaSuccess) cudaGetLastError();
      h->pinned.erase(h->pinned.begin() + k);
      return;
    }
}

void pin_drop_all(Handle* h) {
  while (!h->pinned.empty()) pin_drop(h, h->pinned.back().base);
}

// true when [base, base + bytes) is registered after the call
bool pin_take(Handle* h, const void* base, size_t bytes) {
  if (!h->pin_host || h->device < 0 || bytes < (size_t(4) << 20)) return false;
  if (pinned_is(h, base, bytes)) return true;
  pin_drop(h, base);  // the same address with another extent: a vector that moved back to an old block
  if (cudaHostRegister(const_cast<void*>(base), bytes, cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();  // not fatal: the staged path works for any memory
    return false;
  }
  h->pinned.push_back({base, bytes});
  return true;
}

// room for `extra` more entries; a vector that has to move is unregistered first
template <typename T>
void make_room(Handle* h, std::vector<T>& v, size_t extra) {
  if (v.size() + extra <= v.capacity()) return;
  if (!h->pinned.empty()) pin_drop(h, v.data());
  v.reserve(std::max(v.size() + extra, v.capacity() * 2));
}

// Drop every element inserted at or after element `index` of `family` (global insertion order).
void rollback_to(Handle* h, int family, size_t index) {
  size_t keep[kFamilies] = {0, 0, 0};
  size_t new_journal = 0;
  bool found = false;
  for (size_t j = 0; j < h->journal.size() && !found; ++j) {
    int f = h->journal[j].first;
    size_t cnt = h->journal[j].second;
    if (f == family && index < keep[f] + cnt) {
      size_t take = index - keep[f];
      keep[f] += take;
      h->journal[j].second = take;
      new_journal = take ? j + 1 : j;
      found = true;
    } else {
      keep[f] += cnt;
    }
  }
  if (!found) return;
  if (!h->pinned.empty()) cudaStreamSynchronize(h->stream);  // an upload may still be reading what is dropped here
  h->journal.resize(new_journal);
  for (int f = 0; f < kFamilies; ++f) {
    FamilyHost& fh = h->fh[f];
    size_t n = fh.size();
    if (keep[f] >= n) continue;
    for (size_t i = keep[f]; i < n; ++i) {
      fh.by_number.erase(fh.number[i]);
      uint32_t nd[4] = {fh.conn[0][i], fh.conn[1][i], f == FEMGPU_PLATE ? fh.conn[2][i] : 0u,
                        f == FEMGPU_PLATE ? fh.conn[3][i] : 0u};
      h->nodeset_seen[f].erase(nodeset_hash(f, nd), uint32_t(i), [](uint32_t) { return true; });
    }
    fh.number.resize(keep[f]);
    for (int c = 0; c < kNodesPerElem[f]; ++c) {
      fh.conn[c].resize(keep[f]);
      fh.conn_number[c].resize(keep[f]);
    }
    for (int p = 0; p < kPropsPerElem[f]; ++p) fh.props[p].resize(keep[f]);
    fh.cbase_truncate(keep[f]);
    h->fd[f].uploaded = std::min(h->fd[f].uploaded, keep[f]);
    h->fd[f].validated = std::min(h->fd[f].validated, keep[f]);
  }
  int64_t nc = 0;
  for (int f = 0; f < kFamilies; ++f) nc += int64_t(h->fh[f].size()) * kPairsPerElem[f];
  h->n_contrib = nc;
  invalidate(h);
}

// device validation of everything pending; on error rolls back and fills the reference message
int32_t no_device(Handle* h) {
  return h->fail(FEMGPU_ERR_NO_DEVICE, "this handle was created without a CUDA device (staging only); "
                                       "femgpu has no CPU fallback");
}

int32_t validate_pending(Handle* h, int32_t* family, uint32_t* number, int32_t* code) {
  bool pending = false;
  for (int f = 0; f < kFamilies; ++f) pending |= h->fd[f].validated < h->fh[f].size();
  if (!pending) return 0;
  if (h->device < 0) return no_device(h);
  int32_t st = upload_pending(h);
  if (st) return st;
  st = run_prep(h, /*validate_only=*/true);
  if (st) return st;
  int ef = -1, ec = 0;
  size_t ei = 0;
  st = first_error(h, &ef, &ei, &ec);
  if (st) return st;
  if (ef < 0) {
    for (int f = 0; f < kFamilies; ++f) h->fd[f].validated = h->fh[f].size();
    return 0;
  }
  std::string text = element_error_text(h, ef, ei, ec);
  if (family) *family = ef;
  if (number) *number = h->fh[ef].number[ei];
  if (code) *code = ec;
  rollback_to(h, ef, ei);
  for (int f = 0; f < kFamilies; ++f) h->fd[f].validated = h->fh[f].size();
  return h->fail(ec, text);
}

// A host-side check failed at element k of a batch: the reference would have reported any error of
// an earlier element first, so validate the accepted prefix on the device before answering.
int32_t fail_after_prefix(Handle* h, int32_t code, const std::string& text) {
  if (h->device < 0) return h->fail(code, text);
  int32_t st = validate_pending(h, nullptr, nullptr, nullptr);
  if (st) return st;
  return h->fail(code, text);
}

// With registered staging (pin_take) the accepted part of a large batch goes to the device right away: the copies
// are DMAs queued on the handle's stream, they run under the host checks of the next add_* call.
void upload_early(Handle* h, size_t accepted) {
  if (!h->pin_host || h->device < 0 || accepted < 65536) return;
  if (upload_pending(h)) h->last = Status();  // nothing is lost: the symbolic pass uploads again and reports
}

// Batched FEM::add_truss / add_beam / add_plate: every check of the reference, in the reference's
// order per element (node 1..k exist -> element number unused -> node set unused -> property
// signs), evaluated for the whole batch on all host cores; the batch is accepted up to the first
// failing element.
int32_t add_elements(Handle* h, int family, size_t n, const uint32_t* number,
                     const uint32_t* const* nodes, const double* const* props) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (n == 0) return 0;
  if (!number) return h->fail(FEMGPU_ERR_USAGE, "null element number array");
  const int nn = kNodesPerElem[family], np = kPropsPerElem[family];
  FamilyHost& fh = h->fh[family];
  const size_t start = fh.size();
  if (start + n >= (1u << 26))
    return h->fail(FEMGPU_ERR_LIMIT, "more than 2^26 elements of one family on one device");

  static const bool timing = getenv("FEMGPU_HOST_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[femgpu add %d] %-22s %7.2f ms\n", family, what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  // A. node numbers -> indices (check_node_exist, argument order), in parallel
  std::vector<uint32_t>* idx = h->add_idx;  // scratch kept by the handle: no fresh pages per batch
  for (int c = 0; c < nn; ++c) idx[c].resize(n);
  std::atomic<size_t> i_node(n);
  parallel_chunks(n, 8192, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      bool ok = true;
      for (int c = 0; c < nn; ++c) {
        uint32_t v = 0;
        if (!h->node_by_number.find(nodes[c][i], &v)) ok = false;
        idx[c][i] = v;
      }
      if (!ok) {
        size_t cur = i_node.load();
        while (i < cur && !i_node.compare_exchange_weak(cur, i)) {
        }
        return;  // later elements of this chunk cannot be the first failure
      }
    }
  });
  size_t limit = i_node.load();  // elements at or after this index are never accepted
  lap("node lookup");

  // D. property sign checks of *::create, in parallel (geometry checks run on the device)
  std::atomic<size_t> i_prop(limit);
  parallel_chunks(limit, 8192, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      double pv[11];
      for (int p = 0; p < np; ++p) pv[p] = props[p] ? props[p][i] : NAN;
      if (property_check(family, pv)) {
        size_t cur = i_prop.load();
        while (i < cur && !i_prop.compare_exchange_weak(cur, i)) {
        }
        return;
      }
    }
  });
  lap("property checks");
  // a property failure at i still lets the number / node-set checks of element i run first
  size_t scan_end = std::min(limit, i_prop.load() + 1);

  // B. duplicate element numbers (NumberMap::insert_batch: all cores for large batches)
  const size_t i_num = fh.by_number.insert_batch(number, scan_end, uint32_t(start)), inserted_numbers = i_num;
  scan_end = std::min(scan_end, i_num + 1);
  lap("element numbers");

  // C. duplicate node sets, sharded hash index filled by all cores
  bool degenerate = false;
  std::vector<uint64_t>& hashes = h->add_hash;
  hashes.resize(scan_end);
  parallel_chunks(scan_end, 8192, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) {
      uint32_t nd[4] = {idx[0][i], idx[1][i], nn == 4 ? idx[2][i] : 0u, nn == 4 ? idx[3][i] : 0u};
      hashes[i] = nodeset_hash(family, nd);
    }
  });
  auto node_set_of = [&](uint32_t id, uint32_t out[4]) {
    if (id >= start) {
      size_t i = id - start;
      for (int c = 0; c < 4; ++c) out[c] = c < nn ? idx[c][i] : 0u;
    } else {
      for (int c = 0; c < 4; ++c) out[c] = c < nn ? fh.conn[c][id] : 0u;
    }
  };
  size_t i_set = h->nodeset_seen[family].insert_batch(
      hashes.data(), scan_end, uint32_t(start), [&](uint32_t existing, size_t i) {
        uint32_t a[4], b[4];
        node_set_of(existing, a);
        node_set_of(uint32_t(start + i), b);
        return nodeset_equal(family, a, b);
      }, h->add_items);
  lap("node-set index");
  if (family == FEMGPU_PLATE) {
    // Plate::is_nodes_numbers_same is a subset test; with repeated node numbers in the new element
    // that is not set equality, so such (degenerate) elements are compared by a scan.
    // first element with a repeated node, found on all cores; the scan below only runs for that rare element
    const size_t deg_end = std::min(scan_end, i_set + 1);
    std::atomic<size_t> first_deg(deg_end);
    parallel_chunks(deg_end, 65536, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; ++i) {
        uint32_t v[4] = {idx[0][i], idx[1][i], idx[2][i], idx[3][i]};
        sort4(v);
        if (v[0] == v[1] || v[1] == v[2] || v[2] == v[3]) {
          size_t cur = first_deg.load();
          while (i < cur && !first_deg.compare_exchange_weak(cur, i)) {
          }
          return;
        }
      }
    });
    for (size_t i = first_deg.load(); i < deg_end && !degenerate; ++i) {
      uint32_t v[4] = {idx[0][i], idx[1][i], idx[2][i], idx[3][i]};
      sort4(v);
      if (v[0] == v[1] || v[1] == v[2] || v[2] == v[3]) {
        degenerate = true;
        for (size_t j = 0; j < start + i; ++j) {
          uint32_t o[4];
          node_set_of(uint32_t(j), o);
          bool all = true;
          for (int c = 0; c < 4 && all; ++c) all = (v[c] == o[0] || v[c] == o[1] || v[c] == o[2] || v[c] == o[3]);
          if (all) {
            i_set = std::min(i_set, i);
            break;
          }
        }
      }
    }
  }

  lap("degenerate scan");
  // first failing element and, for it, the first failing check in the reference's order
  size_t e_star = std::min(std::min(limit, i_prop.load()), std::min(i_num, i_set));
  int32_t status = 0;
  std::string text;
  if (e_star < n) {
    const size_t e = e_star;
    if (e == i_node.load()) {
      for (int c = 0; c < nn; ++c) {
        uint32_t v;
        if (!h->node_by_number.find(nodes[c][e], &v)) {
          status = FEMGPU_E_NODE_NOT_EXIST;
          text = "Node with number " + std::to_string(nodes[c][e]) + " does not exist!";
          break;
        }
      }
    } else if (e == i_num) {
      status = FEMGPU_E_ELEMENT_NUMBER_EXISTS;
      text = std::string(family_name(family)) + " element with number " + std::to_string(number[e]) +
             " already exists!";
    } else if (e == i_set) {
      status = FEMGPU_E_ELEMENT_SAME_NODES;
      if (family == FEMGPU_PLATE)
        text = "Plate element with nodes numbers [" + std::to_string(nodes[0][e]) + ", " +
               std::to_string(nodes[1][e]) + ", " + std::to_string(nodes[2][e]) + ", " +
               std::to_string(nodes[3][e]) + "] already exists!";
      else
        text = std::string(family_name(family)) + " element with node number " + std::to_string(nodes[0][e]) +
               " and " + std::to_string(nodes[1][e]) + " already exists!";
    }
  }
  // undo index entries of rejected elements
  const size_t accepted = e_star;
  for (size_t i = accepted; i < inserted_numbers; ++i) fh.by_number.erase(number[i]);
  if (accepted < scan_end) h->nodeset_seen[family].erase_batch(hashes.data(), accepted, scan_end, uint32_t(start));

  // append the accepted prefix: one bulk copy per array, the arrays spread over the host cores
  {
    for (int c = 0; c < nn; ++c) make_room(h, fh.conn[c], accepted);
    for (int p = 0; p < np; ++p) make_room(h, fh.props[p], accepted);
    std::vector<std::function<void()>> jobs;
    jobs.emplace_back([&] { fh.number.insert(fh.number.end(), number, number + accepted); });
    for (int c = 0; c < nn; ++c) {
      jobs.emplace_back([&, c] { fh.conn[c].insert(fh.conn[c].end(), idx[c].begin(), idx[c].begin() + accepted); });
      jobs.emplace_back([&, c] { fh.conn_number[c].insert(fh.conn_number[c].end(), nodes[c], nodes[c] + accepted); });
    }
    for (int p = 0; p < np; ++p)
      jobs.emplace_back([&, p] {
        if (props[p]) fh.props[p].insert(fh.props[p].end(), props[p], props[p] + accepted);
        else fh.props[p].insert(fh.props[p].end(), accepted, NAN);
      });
    const unsigned T = accepted >= 65536 ? std::min<unsigned>(host_threads(), unsigned(jobs.size())) : 1u;
    std::atomic<size_t> next(0);
    parallel_run(T, [&](unsigned, unsigned) {
      for (size_t j = next.fetch_add(1); j < jobs.size(); j = next.fetch_add(1)) jobs[j]();
    });
  }
  lap("appends");
  fh.cbase_append(start, accepted, h->n_contrib, kPairsPerElem[family]);
  h->n_contrib += int64_t(accepted) * kPairsPerElem[family];
  if (accepted) {
    if (!h->journal.empty() && h->journal.back().first == family)
      h->journal.back().second += accepted;
    else
      h->journal.emplace_back(family, accepted);
    invalidate(h);
  }
  if (e_star < n && !status) {
    // property failure: format the message from the caller's arrays (the element was not stored)
    double pv[11];
    for (int p = 0; p < np; ++p) pv[p] = props[p] ? props[p][e_star] : NAN;
    status = property_check(family, pv);
    // element_error_text reads the host arrays: stage the element temporarily
    fh.number.push_back(number[e_star]);
    for (int p = 0; p < np; ++p) make_room(h, fh.props[p], 1);
    for (int p = 0; p < np; ++p) fh.props[p].push_back(pv[p]);
    text = element_error_text(h, family, fh.number.size() - 1, status);
    fh.number.pop_back();
    for (int p = 0; p < np; ++p) fh.props[p].pop_back();
  }
  if (status) return fail_after_prefix(h, status, text);
  upload_early(h, accepted);
  return 0;
}

// Host -> device copy of a pageable host range, stream-ordered. Small ranges go straight through
// cudaMemcpyAsync (the driver stages them); large ones are pipelined through two pinned 16 MB bounce
// buffers filled by several host threads, which is several times faster than the driver's
// single-threaded pageable path. The source may be reused as soon as the call returns.
int32_t h2d_staged(Handle* h, void* dst, const void* src, size_t bytes) {
  if (bytes < (size_t(4) << 20)) {
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return 0;
  }
  for (int b = 0; b < 2; ++b)
    if (!h->pin_buf[b]) {
      FEMGPU_CUDA_CHECK(h, cudaHostAlloc(&h->pin_buf[b], Handle::kPinChunk, cudaHostAllocDefault));
      FEMGPU_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->pin_ev[b], cudaEventDisableTiming));
    }
  const char* s = static_cast<const char*>(src);
  char* d = static_cast<char*>(dst);
  for (size_t off = 0; off < bytes; off += Handle::kPinChunk) {
    const size_t n = std::min(Handle::kPinChunk, bytes - off);
    const int b = h->pin_next;
    h->pin_next ^= 1;
    if (h->pin_busy[b]) FEMGPU_CUDA_CHECK(h, cudaEventSynchronize(h->pin_ev[b]));
    char* stage = static_cast<char*>(h->pin_buf[b]);
    parallel_chunks(n, size_t(2) << 20, [&](size_t lo, size_t hi) { memcpy(stage + lo, s + off + lo, hi - lo); });
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(d + off, stage, n, cudaMemcpyHostToDevice, h->stream));
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->pin_ev[b], h->stream));
    h->pin_busy[b] = true;
  }
  return 0;
}

// entries [from, size) of a staging vector to the same positions of `dst`: a DMA out of the vector when it is
// registered with CUDA (see pin_take), else through the pinned chunks
template <typename T>
int32_t h2d(Handle* h, T* dst, const std::vector<T>& host, size_t from) {
  const size_t bytes = (host.size() - from) * sizeof(T);
  if (pin_take(h, host.data(), host.capacity() * sizeof(T))) {
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(dst + from, host.data() + from, bytes, cudaMemcpyHostToDevice, h->stream));
    return 0;
  }
  return h2d_staged(h, dst + from, host.data() + from, bytes);
}

template <typename T>
int32_t append_to_device(Handle* h, DevBuf<T>& buf, const std::vector<T>& host, size_t from) {
  size_t n = host.size();
  if (n <= from) return 0;
  if (n > buf.cap) {
    // grow: allocate new, re-upload everything (simple; growth is geometric). Stream-ordered: the
    // old buffer is only read by work already queued on h->stream, and cudaFree synchronises.
    DevBuf<T> nb;
    nb.tally = buf.tally;
    nb.stream = buf.stream;
    FEMGPU_CUDA_CHECK(h, nb.reserve(n));
    int32_t st = h2d(h, nb.p, host, 0);
    if (st) return st;
    buf.release();
    buf = nb;
    return 0;
  }
  return h2d(h, buf.p, host, from);
}

}  // namespace

namespace femgpu {

namespace {

__global__ void cbase_fill_kernel(int64_t* __restrict__ out, uint32_t count, int64_t base, int pairs) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = base + int64_t(i) * pairs;
}

// device cbase of the elements [from, size) of family f, run by run (FamilyHost::cbase_runs)
int32_t fill_cbase(Handle* h, int f, size_t from) {
  FamilyHost& fh = h->fh[f];
  FamilyDev& fd = h->fd[f];
  const size_t n = fh.size();
  if (n <= from) return 0;
  if (n > fd.cbase.cap) {  // grow: everything again, like append_to_device
    fd.cbase.release();
    FEMGPU_CUDA_CHECK(h, fd.cbase.reserve(n));
    from = 0;
  }
  for (const auto& r : fh.cbase_runs) {
    const size_t b = std::max(r.start, from), e = r.start + r.count;
    if (b >= e) continue;
    cbase_fill_kernel<<<div_up(e - b, 256), 256, 0, h->stream>>>(fd.cbase.p + b, uint32_t(e - b),
                                                               r.base + int64_t(b - r.start) * kPairsPerElem[f], kPairsPerElem[f]);
    h->launches++;
  }
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  return 0;
}

}  // namespace

int32_t upload_pending(Handle* h) {
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  int32_t st;
  if ((st = append_to_device(h, h->d_x, h->nx, h->nodes_uploaded))) return st;
  if ((st = append_to_device(h, h->d_y, h->ny, h->nodes_uploaded))) return st;
  if ((st = append_to_device(h, h->d_z, h->nz, h->nodes_uploaded))) return st;
  h->nodes_uploaded = h->n_nodes();
  for (int f = 0; f < kFamilies; ++f) {
    FamilyHost& fh = h->fh[f];
    FamilyDev& fd = h->fd[f];
    for (int c = 0; c < kNodesPerElem[f]; ++c)
      if ((st = append_to_device(h, fd.conn[c], fh.conn[c], fd.uploaded))) return st;
    for (int p = 0; p < kPropsPerElem[f]; ++p)
      if ((st = append_to_device(h, fd.props[p], fh.props[p], fd.uploaded))) return st;
    if ((st = fill_cbase(h, f, fd.uploaded))) return st;
    fd.uploaded = fh.size();
  }
  return 0;
}

// The library's private stream-ordered pool, one per device, created on first use and kept for the life of the
// process (freed memory stays mapped: release threshold lifted). The default pool of the device is not touched.
cudaError_t pool_malloc(void** p, size_t bytes, cudaStream_t s) {
  static cudaMemPool_t pools[64] = {};
  static std::mutex mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaMallocAsync(p, bytes, s);
  cudaMemPool_t pool;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[dev]) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      if ((e = cudaMemPoolCreate(&pools[dev], &props)) != cudaSuccess) return e;
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool = pools[dev];
  }
  return cudaMallocFromPoolAsync(p, bytes, pool, s);
}

}  // namespace femgpu

extern "C" {

int32_t femgpu_create(femgpu_t** out, double rel_tol, double abs_tol, uint32_t nodes_number,
                      int32_t device) {
  if (!out) return FEMGPU_ERR_USAGE;
  *out = nullptr;
  if (device == FEMGPU_DEVICE_NONE) {
    // staging-only handle: host bookkeeping (numbering, duplicate checks, error texts) works,
    // every call that needs the GPU answers FEMGPU_ERR_NO_DEVICE. Used by the CPU-side tests.
    femgpu_t* h = new femgpu_t();
    h->rel_tol = rel_tol;
    h->abs_tol = abs_tol;
    h->nodes_number = nodes_number;
    h->device = FEMGPU_DEVICE_NONE;
    *out = h;
    return 0;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_status.code = FEMGPU_ERR_NO_DEVICE;
    g_create_status.text = std::string("no CUDA device available (") +
                           (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                           "); femgpu has no CPU fallback";
    return FEMGPU_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) {
    g_create_status.code = FEMGPU_ERR_USAGE;
    g_create_status.text = "device ordinal out of range";
    return FEMGPU_ERR_USAGE;
  }
  femgpu_t* h = new femgpu_t();
  h->rel_tol = rel_tol;
  h->abs_tol = abs_tol;
  h->nodes_number = nodes_number;
  h->device = device;
  // the handle's stream gets the highest priority: when the element-record kernels of the next slab range share the
  // device with the assembly kernel (femgpu_numeric), freed SM resources go to the assembly CTAs first
  int prio_least = 0, prio_greatest = 0;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) {
    g_create_status.code = FEMGPU_ERR_CUDA;
    g_create_status.text = "could not create a CUDA stream";
    delete h;
    return FEMGPU_ERR_CUDA;
  }
  for (auto& q : h->ev)
    for (auto& ev : q) cudaEventCreate(&ev);
  if (const char* e = getenv("FEMGPU_PIN_HOST")) h->pin_host = atoi(e) == 1;  // 1: from the first upload on, not the first reset
  *out = h;
  return 0;
}

static void free_device(femgpu_t* h) {
  if (h->device < 0) return;
  cudaSetDevice(h->device);
  h->d_x.release(); h->d_y.release(); h->d_z.release();
  for (auto& f : h->fd) {
    for (auto& c : f.conn) c.release();
    for (auto& p : f.props) p.release();
    f.cbase.release(); f.rec.release(); f.err.release();
    f.uploaded = f.validated = 0;
  }
  h->nz_row_ptr.release(); h->nz_col.release(); h->nz_val.release(); h->nz_valid = false;
  h->trig_table.release();
  h->blk_key.release(); h->blk_full.release(); h->blk_cptr.release(); h->contrib.release();
  h->blk_meta.release(); h->blk_order.release(); h->items.release(); h->items_c.release(); h->elist.release(); h->elist_compact.release(); h->node_blk_ptr.release(); h->node_base.release();
  h->node_len.release(); h->blk_off.release(); h->slabs.release(); h->row_ptr.release();
  h->col_idx.release(); h->values.release(); h->scratch.release(); h->d_flag.release();
  h->dist.send_buf.release(); h->dist.recv_buf.release(); h->dist.recv_dst_block.release();
  h->dist.recv_full.release(); h->dist.remote_keys.release();
  sep_release(h);
  sol_release(h);
}

int32_t femgpu_reset(femgpu_t* h, uint32_t nodes_number) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device >= 0) cudaStreamSynchronize(h->stream);
  if (h->device >= 0 && !h->pin_host) {  // a re-used instance: its staging is worth registering (FEMGPU_PIN_HOST=0: never)
    const char* e = getenv("FEMGPU_PIN_HOST");
    h->pin_host = !(e && atoi(e) == 0);
  }
  // FEM::reset of a re-used instance: the device buffers stay with the handle (every one of them is rewritten before
  // it is read again — uploads restart at element 0, the symbolic products are rebuilt), like the host staging does.
  // Handing ~20 GB back to the stream-ordered pool and asking for it again cost 0.04 s on most boxes and 0.7-0.9 s
  // on some (profiles/README.md); femgpu_destroy frees everything. FEMGPU_RESET_FREES=1 restores the old behaviour.
  static const bool reset_frees = getenv("FEMGPU_RESET_FREES") != nullptr;
  if (reset_frees) {
    free_device(h);
  } else if (h->device >= 0) {
    for (auto& f : h->fd) f.uploaded = f.validated = 0;
    sol_invalidate(h);
  }
  h->nodes_number = nodes_number;
  h->node_index_base = 0;
  h->node_number.clear(); h->nx.clear(); h->ny.clear(); h->nz.clear();
  h->node_by_number.clear(); h->node_by_xyz.clear();
  h->nodes_uploaded = 0;
  for (auto& f : h->fh) f.clear();
  for (auto& x : h->nodeset_seen) x.clear();
  h->n_contrib = 0;
  h->journal.clear();
  bc_clear(h);
  invalidate(h);
  h->n_rows = h->nnz = 0;
  h->n_blocks = h->n_slabs = 0;
  return 0;
}

void femgpu_destroy(femgpu_t* h) {
  if (!h) return;
  if (h->device < 0) {
    delete h;
    return;
  }
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  pin_drop_all(h);
  dist_destroy(h);
  free_device(h);
  for (auto& q : h->ev)
    for (auto& ev : q)
      if (ev) cudaEventDestroy(ev);
  for (int b = 0; b < 2; ++b) {
    if (h->side_stream[b]) cudaStreamDestroy(h->side_stream[b]);
    if (h->join_ev[b]) cudaEventDestroy(h->join_ev[b]);
  }
  if (h->fork_ev) cudaEventDestroy(h->fork_ev);
  if (h->prep_stream) cudaStreamDestroy(h->prep_stream);
  if (h->xchg_stream) cudaStreamDestroy(h->xchg_stream);
  if (h->ghost_ev) cudaEventDestroy(h->ghost_ev);
  if (h->pack_ev) cudaEventDestroy(h->pack_ev);
  for (auto& e : h->range_ev)
    if (e) cudaEventDestroy(e);
  for (auto& q : h->range_t0)
    for (auto& e : q)
      if (e) cudaEventDestroy(e);
  for (auto& q : h->range_t1)
    for (auto& e : q)
      if (e) cudaEventDestroy(e);
  for (int b = 0; b < 2; ++b) {
    if (h->pin_ev[b]) cudaEventDestroy(h->pin_ev[b]);
    if (h->pin_buf[b]) cudaFreeHost(h->pin_buf[b]);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* femgpu_last_error(const femgpu_t* h) {
  if (!h) return g_create_status.text.c_str();
  return h->last.text.c_str();
}

int32_t femgpu_add_nodes(femgpu_t* h, size_t n, const uint32_t* number, const double* x,
                         const double* y, const double* z) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (n && (!number || !x || !y || !z)) return h->fail(FEMGPU_ERR_USAGE, "null node array");
  if (n == 0) return 0;
  const size_t n0 = h->n_nodes();
  const size_t g0 = size_t(h->node_index_base) + n0;  // global insertion index of the batch's first node
  // methods_for_node_data_handle.rs:42-64, per node: limit, then number, (index,) coordinates
  const size_t room = h->nodes_number > g0 ? size_t(h->nodes_number) - g0 : 0;
  const size_t i_limit = std::min(n, room);
  size_t scan_end = std::min(n, i_limit + 0);
  static const bool timing = getenv("FEMGPU_HOST_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[femgpu add nodes] %-22s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  // duplicate numbers (NumberMap::insert_batch)
  const size_t i_num = h->node_by_number.insert_batch(number, scan_end, uint32_t(g0)), inserted = i_num;
  scan_end = std::min(scan_end, i_num + 1);
  lap("node numbers");
  // duplicate coordinates: sharded hash index, all cores (NaN never compares equal)
  std::vector<uint64_t>& hashes = h->add_hash;
  hashes.resize(scan_end);
  parallel_chunks(scan_end, 8192, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; ++i) hashes[i] = xyz_hash(x[i], y[i], z[i]);
  });
  size_t i_xyz = h->node_by_xyz.insert_batch(hashes.data(), scan_end, uint32_t(n0), [&](uint32_t existing, size_t i) {
    double ex, ey, ez;
    if (existing >= n0) {
      ex = x[existing - n0]; ey = y[existing - n0]; ez = z[existing - n0];
    } else {
      ex = h->nx[existing]; ey = h->ny[existing]; ez = h->nz[existing];
    }
    return ex == x[i] && ey == y[i] && ez == z[i];
  }, h->add_items);
  lap("coordinate index");
  const size_t accepted = std::min(std::min(i_limit, i_num), i_xyz);
  for (size_t i = accepted; i < inserted; ++i) h->node_by_number.erase(number[i]);
  if (accepted < scan_end) h->node_by_xyz.erase_batch(hashes.data(), accepted, scan_end, uint32_t(n0));
  make_room(h, h->nx, accepted);
  make_room(h, h->ny, accepted);
  make_room(h, h->nz, accepted);
  parallel_run(accepted >= 65536 ? std::min(4u, host_threads()) : 1u, [&](unsigned t, unsigned nt) {
    for (unsigned j = t; j < 4u; j += nt) {
      if (j == 0) h->node_number.insert(h->node_number.end(), number, number + accepted);
      if (j == 1) h->nx.insert(h->nx.end(), x, x + accepted);
      if (j == 2) h->ny.insert(h->ny.end(), y, y + accepted);
      if (j == 3) h->nz.insert(h->nz.end(), z, z + accepted);
    }
  });
  lap("appends");
  if (accepted) invalidate(h);
  if (accepted == n) upload_early(h, accepted);
  if (accepted < n) {
    const size_t e = accepted;
    if (e == i_limit)
      return h->fail(FEMGPU_E_NODE_LIMIT,
                     "Nodes number could not be greater than " + std::to_string(h->nodes_number) + "!");
    if (e == i_num)
      return h->fail(FEMGPU_E_NODE_NUMBER_EXISTS, "Node with number " + std::to_string(number[e]) + " already exists!");
    return h->fail(FEMGPU_E_NODE_COORDINATES_EXIST,
                   "Node with coordinates x: " + rust_debug_f64(x[e]) + ", y: " + rust_debug_f64(y[e]) +
                       ", z: " + rust_debug_f64(z[e]) + " already exists!");
  }
  return 0;
}

int32_t femgpu_add_truss(femgpu_t* h, size_t n, const uint32_t* number, const uint32_t* node_1,
                         const uint32_t* node_2, const double* young_modulus, const double* area,
                         const double* area_2) {
  const uint32_t* nodes[4] = {node_1, node_2, nullptr, nullptr};
  const double* props[11] = {young_modulus, area, area_2};
  if (h && n && (!node_1 || !node_2 || !young_modulus || !area))
    return h->fail(FEMGPU_ERR_USAGE, "null truss array");
  return add_elements(h, FEMGPU_TRUSS, n, number, nodes, props);
}

int32_t femgpu_add_beam(femgpu_t* h, size_t n, const uint32_t* number, const uint32_t* node_1,
                        const uint32_t* node_2, const double* young_modulus,
                        const double* poisson_ratio, const double* area, const double* i11,
                        const double* i22, const double* i12, const double* it,
                        const double* shear_factor, const double* local_axis_1) {
  if (h && n && (!node_1 || !node_2 || !young_modulus || !poisson_ratio || !area || !i11 || !i22 ||
                 !i12 || !it || !shear_factor || !local_axis_1))
    return h->fail(FEMGPU_ERR_USAGE, "null beam array");
  const uint32_t* nodes[4] = {node_1, node_2, nullptr, nullptr};
  const double* props[11] = {young_modulus, poisson_ratio, area, i11, i22, i12, it, shear_factor,
                             local_axis_1, local_axis_1 ? local_axis_1 + n : nullptr,
                             local_axis_1 ? local_axis_1 + 2 * n : nullptr};
  return add_elements(h, FEMGPU_BEAM, n, number, nodes, props);
}

int32_t femgpu_add_plate(femgpu_t* h, size_t n, const uint32_t* number, const uint32_t* node_1,
                         const uint32_t* node_2, const uint32_t* node_3, const uint32_t* node_4,
                         const double* young_modulus, const double* poisson_ratio,
                         const double* thickness, const double* shear_factor) {
  if (h && n && (!node_1 || !node_2 || !node_3 || !node_4 || !young_modulus || !poisson_ratio ||
                 !thickness || !shear_factor))
    return h->fail(FEMGPU_ERR_USAGE, "null plate array");
  const uint32_t* nodes[4] = {node_1, node_2, node_3, node_4};
  const double* props[11] = {young_modulus, poisson_ratio, thickness, shear_factor};
  return add_elements(h, FEMGPU_PLATE, n, number, nodes, props);
}

int32_t femgpu_validate(femgpu_t* h, int32_t* family, uint32_t* number, int32_t* code) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  return validate_pending(h, family, number, code);
}

int32_t femgpu_counts(const femgpu_t* h, uint64_t* nodes, uint64_t* truss, uint64_t* beam,
                      uint64_t* plate) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (nodes) *nodes = h->n_nodes();
  if (truss) *truss = h->fh[0].size();
  if (beam) *beam = h->fh[1].size();
  if (plate) *plate = h->fh[2].size();
  return 0;
}

int32_t femgpu_get_numbers(const femgpu_t* h, int32_t family, uint32_t* out) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (family < -1 || family >= femgpu::kFamilies) return h->fail(FEMGPU_ERR_USAGE, "family must be -1 (nodes), 0, 1 or 2");
  const std::vector<uint32_t>& v = family < 0 ? h->node_number : h->fh[family].number;
  if (out && !v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(uint32_t));
  return 0;
}

int32_t femgpu_symbolic(femgpu_t* h, int64_t* n_rows, int64_t* nnz) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  const bool timing = getenv("FEMGPU_SYM_TIMING") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(h->stream);
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[femgpu symbolic] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t0).count());
    t0 = now;
  };
  int32_t st = validate_pending(h, nullptr, nullptr, nullptr);
  if (st) return st;
  lap("upload + device validation");
  if (!h->symbolic_valid) {
    h->values_valid = false;  // the value array is re-laid-out: nothing assembled yet on the new pattern
    if ((st = upload_pending(h))) return st;
    if ((st = run_symbolic(h))) return st;
    h->symbolic_valid = true;
    lap("run_symbolic total");
  }
  if (n_rows) *n_rows = h->n_rows;
  if (nnz) *nnz = h->nnz;
  return 0;
}

int32_t femgpu_numeric(femgpu_t* h) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  if (!h->symbolic_valid) return h->fail(FEMGPU_ERR_USAGE, "femgpu_numeric before femgpu_symbolic");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  const uint32_t slot = uint32_t(h->n_numeric % Handle::kEvRing);
  cudaEvent_t* ev = h->ev[slot];
  FEMGPU_CUDA_CHECK(h, cudaEventRecord(ev[0], h->stream));
  int32_t st = 0;
  const int R = h->n_ranges;
  h->range_count[slot] = 0;
  auto range_events = [&]() -> int32_t {
    if (h->range_t0[0][0]) return 0;
    for (auto& q : h->range_t0)
      for (auto& e : q) FEMGPU_CUDA_CHECK(h, cudaEventCreate(&e));
    for (auto& q : h->range_t1)
      for (auto& e : q) FEMGPU_CUDA_CHECK(h, cudaEventCreate(&e));
    return 0;
  };
  uint32_t g0 = 0;
  if (dist_ghost_first(h, &g0)) {
    // Multi-GPU with peer windows: ghost slabs first, their blocks leave on a second stream while the rest of the
    // matrix is assembled; the partials of the lower neighbour are added at the end (dist.cu).
    if (!h->xchg_stream) {
      FEMGPU_CUDA_CHECK(h, cudaStreamCreateWithFlags(&h->xchg_stream, cudaStreamNonBlocking));
      FEMGPU_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->ghost_ev, cudaEventDisableTiming));
      FEMGPU_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->pack_ev, cudaEventDisableTiming));
    }
    if ((st = range_events())) return st;
    if ((st = run_prep(h, /*validate_only=*/false))) return st;
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(ev[1], h->stream));
    if ((st = dist_begin_pass(h))) return st;
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->range_t0[slot][0], h->stream));
    if ((st = run_assembly(h, g0, h->n_slabs))) return st;  // (the previous pass's pack was joined at its end)
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->range_t1[slot][0], h->stream));
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->ghost_ev, h->stream));
    FEMGPU_CUDA_CHECK(h, cudaStreamWaitEvent(h->xchg_stream, h->ghost_ev, 0));
    if ((st = dist_pack(h, h->xchg_stream))) return st;
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->pack_ev, h->xchg_stream));
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->range_t0[slot][1], h->stream));
    if ((st = run_assembly(h, 0, g0))) return st;
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->range_t1[slot][1], h->stream));
    h->range_count[slot] = 2;
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(ev[2], h->stream));
    if ((st = dist_apply(h))) return st;
    // join: a synchronised handle has sent its blocks, and the next pass may overwrite the ghost rows
    FEMGPU_CUDA_CHECK(h, cudaStreamWaitEvent(h->stream, h->pack_ev, 0));
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(ev[3], h->stream));
    h->n_numeric++;
    h->values_valid = true;
    return 0;
  }
  if (R <= 1) {
    if ((st = run_prep(h, /*validate_only=*/false))) return st;
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(ev[1], h->stream));
    if ((st = run_assembly(h, 0, h->n_slabs))) return st;
  } else {
    // Software pipeline over slab ranges: the element records range r + 1 needs are computed on a low-priority stream
    // while the (high-priority) assembly kernel works on range r; only range 0's records are on the critical path.
    if (!h->prep_stream) {
      int least = 0, greatest = 0;
      FEMGPU_CUDA_CHECK(h, cudaDeviceGetStreamPriorityRange(&least, &greatest));
      FEMGPU_CUDA_CHECK(h, cudaStreamCreateWithPriority(&h->prep_stream, cudaStreamNonBlocking, least));
      if (!h->fork_ev) FEMGPU_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
      for (auto& e : h->range_ev) FEMGPU_CUDA_CHECK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if ((st = range_events())) return st;
    FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->fork_ev, h->stream));  // after the previous pass: its kernels read the records
    FEMGPU_CUDA_CHECK(h, cudaStreamWaitEvent(h->prep_stream, h->fork_ev, 0));
    auto prep = [&](int r) -> int32_t {
      int32_t e = run_prep_range(h, r, h->prep_stream);
      if (e) return e;
      FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->range_ev[r], h->prep_stream));
      return 0;
    };
    if ((st = prep(0))) return st;
    for (int r = 0; r < R; ++r) {
      if (r + 1 < R && (st = prep(r + 1))) return st;
      FEMGPU_CUDA_CHECK(h, cudaStreamWaitEvent(h->stream, h->range_ev[r], 0));
      if (r == 0) FEMGPU_CUDA_CHECK(h, cudaEventRecord(ev[1], h->stream));
      FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->range_t0[slot][r], h->stream));
      if ((st = run_assembly(h, r ? h->range_slab_end[r - 1] : 0u, h->range_slab_end[r]))) return st;
      FEMGPU_CUDA_CHECK(h, cudaEventRecord(h->range_t1[slot][r], h->stream));
    }
    h->range_count[slot] = R;
  }
  if ((st = run_assembly_unstaged(h))) return st;
  FEMGPU_CUDA_CHECK(h, cudaEventRecord(ev[2], h->stream));
  if (h->dist.enabled && (st = dist_numeric_exchange(h))) return st;
  FEMGPU_CUDA_CHECK(h, cudaEventRecord(ev[3], h->stream));
  h->n_numeric++;
  h->values_valid = true;
  return 0;
}

int32_t femgpu_synchronize(femgpu_t* h) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return dist_check(h);
}

int32_t femgpu_assemble(femgpu_t* h, int64_t* n_rows, int64_t* nnz) {
  int32_t st = femgpu_symbolic(h, n_rows, nnz);
  if (st) return st;
  if ((st = femgpu_numeric(h))) return st;
  return femgpu_synchronize(h);
}

int32_t femgpu_get_csr(femgpu_t* h, int64_t* row_ptr, int32_t* col_idx, double* values) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  if (!h->symbolic_valid) return h->fail(FEMGPU_ERR_USAGE, "no assembled matrix");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  if (row_ptr)
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(row_ptr, h->row_ptr.p, size_t(h->n_rows + 1) * 8,
                                         cudaMemcpyDeviceToHost, h->stream));
  if (col_idx && h->nnz)
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(col_idx, h->col_idx.p, size_t(h->nnz) * 4,
                                         cudaMemcpyDeviceToHost, h->stream));
  if (values && h->nnz)
    FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(values, h->values.p, size_t(h->nnz) * 8,
                                         cudaMemcpyDeviceToHost, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return dist_check(h);
}

int32_t femgpu_get_csr_device(femgpu_t* h, const int64_t** row_ptr, const int32_t** col_idx,
                              const double** values, int64_t* row_begin, int64_t* row_end) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  if (!h->symbolic_valid) return h->fail(FEMGPU_ERR_USAGE, "no assembled matrix");
  if (row_ptr) *row_ptr = h->row_ptr.p;
  if (col_idx) *col_idx = h->col_idx.p;
  if (values) *values = h->values.p;
  int64_t rb = 0, re = h->n_rows;
  if (h->dist.enabled && h->dist.ownership_set) {
    rb = 6 * int64_t(h->dist.own_begin);
    re = 6 * int64_t(h->dist.own_end);
  }
  if (row_begin) *row_begin = rb;
  if (row_end) *row_end = re;
  return 0;
}

int32_t femgpu_get_nonzero_coo(femgpu_t* h, int64_t* count, int64_t* rows, int64_t* cols,
                               double* values) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  if (!h->symbolic_valid) return h->fail(FEMGPU_ERR_USAGE, "no assembled matrix");
  return nonzero_coo(h, count, rows, cols, values);
}

int32_t femgpu_get_nonzero_csr(femgpu_t* h, int64_t* count, int64_t* row_ptr, int32_t* col_idx, double* values) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  if (!h->symbolic_valid || !h->values_valid) return h->fail(FEMGPU_ERR_USAGE, "no assembled matrix");
  int32_t st = nonzero_csr(h, count, row_ptr, col_idx, values);
  if (st) return st;
  return dist_check(h);
}

static int32_t find_element(femgpu_t* h, int32_t family, uint32_t number, size_t* index) {
  if (family < 0 || family >= kFamilies) return h->fail(FEMGPU_ERR_USAGE, "bad family");
  uint32_t idx;
  if (!h->fh[family].by_number.find(number, &idx))
    // check_*_element_exist (methods_for_truss_data_handle.rs:130-135 and siblings)
    return h->fail(FEMGPU_E_ELEMENT_NOT_EXIST, std::string(family_name(family)) + " element with number " +
                                                   std::to_string(number) + " does not exist!");
  *index = idx;
  return 0;
}

int32_t femgpu_rotation_elements(femgpu_t* h, int32_t family, uint32_t number, double out[9]) {
  if (!h || !out) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  size_t idx;
  int32_t st = find_element(h, family, number, &idx);
  if (st) return st;
  if ((st = validate_pending(h, nullptr, nullptr, nullptr))) return st;
  if ((st = upload_pending(h))) return st;
  return element_rotation(h, family, idx, out);
}

int32_t femgpu_element_matrix(femgpu_t* h, int32_t family, uint32_t number, double* out) {
  if (!h || !out) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  size_t idx;
  int32_t st = find_element(h, family, number, &idx);
  if (st) return st;
  if ((st = validate_pending(h, nullptr, nullptr, nullptr))) return st;
  if ((st = upload_pending(h))) return st;
  return element_matrix(h, family, idx, out);
}

int32_t femgpu_element_slots(femgpu_t* h, int32_t family, uint32_t number, int64_t* out) {
  if (!h || !out) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  size_t idx;
  int32_t st = find_element(h, family, number, &idx);
  if (st) return st;
  if (!h->symbolic_valid) return h->fail(FEMGPU_ERR_USAGE, "femgpu_element_slots before femgpu_symbolic");
  return element_slots(h, family, idx, out);
}

int32_t femgpu_launch_count(femgpu_t* h, int32_t reset, uint64_t* launches) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (launches) *launches = h->launches;
  if (reset) h->launches = 0;
  return 0;
}

int32_t femgpu_last_numeric_ms(femgpu_t* h, float out[4]) {
  if (!h || !out) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  return femgpu_numeric_ms_history(h, 0, out);
}

int32_t femgpu_numeric_ms_history(femgpu_t* h, uint32_t passes_back, float out[4]) {
  if (!h || !out) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  if (passes_back >= Handle::kEvRing || passes_back >= h->n_numeric)
    return h->fail(FEMGPU_ERR_USAGE, "no such numeric pass in the timing history");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  cudaEvent_t* ev = h->ev[(h->n_numeric - 1 - passes_back) % Handle::kEvRing];
  FEMGPU_CUDA_CHECK(h, cudaEventSynchronize(ev[3]));
  FEMGPU_CUDA_CHECK(h, cudaEventElapsedTime(&out[0], ev[0], ev[3]));
  FEMGPU_CUDA_CHECK(h, cudaEventElapsedTime(&out[1], ev[0], ev[1]));
  FEMGPU_CUDA_CHECK(h, cudaEventElapsedTime(&out[2], ev[1], ev[2]));
  FEMGPU_CUDA_CHECK(h, cudaEventElapsedTime(&out[3], ev[2], ev[3]));
  return 0;
}

int32_t femgpu_numeric_kernel_ms(femgpu_t* h, uint32_t passes_back, float* assemble_ms, int32_t* launches) {
  if (!h || !assemble_ms) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return no_device(h);
  if (passes_back >= Handle::kEvRing || passes_back >= h->n_numeric)
    return h->fail(FEMGPU_ERR_USAGE, "no such numeric pass in the timing history");
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  const uint32_t slot = uint32_t((h->n_numeric - 1 - passes_back) % Handle::kEvRing);
  cudaEvent_t* ev = h->ev[slot];
  FEMGPU_CUDA_CHECK(h, cudaEventSynchronize(ev[3]));
  const int R = h->range_count[slot];
  float total = 0.f;
  if (R == 0) {
    FEMGPU_CUDA_CHECK(h, cudaEventElapsedTime(&total, ev[1], ev[2]));
  } else {
    for (int r = 0; r < R; ++r) {
      float ms = 0.f;
      FEMGPU_CUDA_CHECK(h, cudaEventElapsedTime(&ms, h->range_t0[slot][r], h->range_t1[slot][r]));
      total += ms;
    }
  }
  *assemble_ms = total;
  if (launches) *launches = R ? R : 1;
  return 0;
}

int32_t femgpu_device_bytes(const femgpu_t* h, uint64_t* bytes) {
  if (!h || !bytes) return FEMGPU_ERR_USAGE;
  *bytes = h->dev_bytes;
  return 0;
}

int32_t femgpu_stream(femgpu_t* h, void** stream) {
  if (!h || !stream) return FEMGPU_ERR_USAGE;
  *stream = (void*)h->stream;
  return 0;
}

}  // extern "C"
