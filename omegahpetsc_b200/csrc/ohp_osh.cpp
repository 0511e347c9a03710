/*
 * ohp_osh.cpp -- reader for Omega_h `.osh` mesh directories (binary format version 9).
 *
 * Replaces Omega_h::binary::read(path, comm, &mesh) (ex12.cpp:813,818), Mesh::ask_elem_verts()
 * (ex12.cpp:591,634,784), mesh.coords() and mesh.get_array<T>(dim, name) (ex12.cpp:593-594,
 * 627-628,703-705).  Omega_h is an un-vendored dependency of the reference (README.md:151,547:
 * tag v9.32.1); the stream layout below is its published on-disk format, anchored on the
 * reference's own fixture files which must decode to the last byte.
 *
 * Directory layout: text files `nparts` and `version`, one stream `<part>.osh` per part.
 * Stream (little endian): magic a1 1a | i8 compressed | i8 family | i8 dim | i32 comm_size |
 * i32 comm_rank | i8 parting | i32 nghost | i8 have_hints [i32 naxes, naxes x 3 f64] | i32 nverts |
 * for d = 1..dim: array<i32> down(d -> d-1), and for d > 1 array<i8> alignment codes |
 * for d = 0..dim: i32 ntags, each tag = (i32 len, name, i8 ncomps, i8 type, array)
 *                 [+ owner ranks/idxs arrays when comm_size > 1] |
 * i32 nsets (class sets) | i8 has-parents.
 * array = i32 n, then (compressed) i64 zbytes + zlib stream, or the raw n items.
 */
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <map>
#include <string>
#include <vector>

#include "ohp_internal.h"

namespace {

struct Tag {
  int ncomps, type;
  int64_t count; /* number of scalars */
  std::vector<uint8_t> bytes;
};

int type_size(int t) { return t == 0 ? 1 : t == 2 ? 4 : t == 3 ? 8 : t == 5 ? 8 : -1; }

struct Cursor {
  const uint8_t *p, *end;
  bool ok = true;
  template <class T> T take() {
    T v{};
    if (p + sizeof(T) > end) { ok = false; return v; }
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  bool array(int item_size, bool compressed, std::vector<uint8_t> &out, int64_t *n_items) {
    const int32_t n = take<int32_t>();
    if (!ok || n < 0) return false;
    const size_t nbytes = (size_t)n * item_size;
    out.resize(nbytes);
    if (compressed) {
      const int64_t zb = take<int64_t>();
      if (!ok || zb < 0 || p + zb > end) return false;
      uLongf dest = (uLongf)nbytes;
      if (nbytes == 0) {
        uint8_t dummy[8]; dest = 8;
        uncompress(dummy, &dest, p, (uLong)zb);
      } else {
        const int rc = uncompress(out.data(), &dest, p, (uLong)zb);
        if (rc != Z_OK || dest != nbytes) return false;
      }
      p += zb;
    } else {
      if (p + nbytes > end) return false;
      memcpy(out.data(), p, nbytes);
      p += nbytes;
    }
    *n_items = n;
    return true;
  }
};

bool read_int_file(const std::string &path, int *v) {
  FILE *f = fopen(path.c_str(), "r");
  if (!f) return false;
  const int k = fscanf(f, "%d", v);
  fclose(f);
  return k == 1;
}

}  // namespace

struct ohp_osh_s {
  int version = 0, nparts = 0;
  int family = 0, dim = 0, comm_size = 0, comm_rank = 0, parting = 0, nghost = 0;
  int32_t nverts = 0;
  std::vector<int32_t> down[4];
  std::vector<int8_t> codes[4];
  std::map<std::string, Tag> tags[4];
  std::vector<int32_t> owner_ranks[4], owner_idxs[4];
  std::vector<int32_t> elem_verts;
  int32_t nents[4] = {0, 0, 0, 0};
};

/* vertex of edge-use `which` (0/1) of a triangle side given the alignment code:
 * code bit 0 = flip, bits 1-2 = rotation, bits 3+ = which_down (Omega_h_align.hpp).
 * For an edge only rotation 0/1 occurs: rotation 1 means the triangle traverses the edge
 * from its second to its first vertex. */
static int derive_elem_verts(ohp_osh_s &m) {
  const int dim = m.dim;
  const std::vector<int32_t> &ev = m.down[1];
  const int64_t nedges = (int64_t)ev.size() / 2;
  if (dim == 1) { m.elem_verts = ev; return 0; }
  const std::vector<int32_t> &te = m.down[2];
  const std::vector<int8_t> &tc = m.codes[2];
  const int64_t ntri = (int64_t)te.size() / 3;
  if ((int64_t)tc.size() != ntri * 3) OHP_FAIL(OHP_ERR_FILE_UNEXPECTED, "osh: tri->edge codes length mismatch");
  std::vector<int32_t> tri((size_t)ntri * 3);
  for (int64_t t = 0; t < ntri; ++t)
    for (int i = 0; i < 3; ++i) {
      const int32_t e = te[t * 3 + i];
      if (e < 0 || e >= nedges) OHP_FAIL(OHP_ERR_FILE_UNEXPECTED, "osh: triangle %lld references edge %d", (long long)t, e);
      const int rot = (tc[t * 3 + i] >> 1) & 3;
      if (rot > 1) OHP_FAIL(OHP_ERR_FILE_UNEXPECTED, "osh: edge rotation %d in tri->edge codes", rot);
      tri[t * 3 + i] = ev[(size_t)e * 2 + rot];
    }
  if (dim == 2) { m.elem_verts.swap(tri); return 0; }
  /* tets: faces 0 = (0,2,1), 1 = (0,1,3) of the simplex down template; vertices 0..2 from the
   * aligned face 0, vertex 3 from the aligned face 1 */
  const std::vector<int32_t> &tt = m.down[3];
  const std::vector<int8_t> &cc = m.codes[3];
  const int64_t ntet = (int64_t)tt.size() / 4;
  if ((int64_t)cc.size() != ntet * 4) OHP_FAIL(OHP_ERR_FILE_UNEXPECTED, "osh: tet->tri codes length mismatch");
  m.elem_verts.resize((size_t)ntet * 4);
  auto aligned = [&](int64_t t, int col, int32_t out[3]) {
    const int32_t f = tt[t * 4 + col];
    const int code = cc[t * 4 + col];
    const int flip = code & 1, r = (code >> 1) & 3;
    for (int k = 0; k < 3; ++k) out[k] = tri[(size_t)f * 3 + ((k + 3 - r) % 3)];
    if (flip) std::swap(out[1], out[2]);
  };
  for (int64_t t = 0; t < ntet; ++t) {
    for (int col = 0; col < 2; ++col)
      if (tt[t * 4 + col] < 0 || tt[t * 4 + col] >= ntri)
        OHP_FAIL(OHP_ERR_FILE_UNEXPECTED, "osh: tet %lld references face %d", (long long)t, tt[t * 4 + col]);
    int32_t f0[3], f1[3];
    aligned(t, 0, f0);
    aligned(t, 1, f1);
    m.elem_verts[t * 4 + 0] = f0[0];
    m.elem_verts[t * 4 + 1] = f0[2];
    m.elem_verts[t * 4 + 2] = f0[1];
    m.elem_verts[t * 4 + 3] = f1[2];
  }
  return 0;
}

extern "C" int ohp_osh_read(const char *dir, int part, ohp_osh *out) {
  if (!dir || !out) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_osh_read: null argument");
  std::string d(dir);
  while (d.size() > 1 && d.back() == '/') d.pop_back();
  ohp_osh_s *m = new ohp_osh_s();
  if (!read_int_file(d + "/version", &m->version) || !read_int_file(d + "/nparts", &m->nparts)) {
    delete m;
    OHP_FAIL(OHP_ERR_FILE_OPEN, "ohp_osh_read: %s is not an Omega_h mesh directory (missing version/nparts)", dir);
  }
  if (m->version < 7 || m->version > 10) {
    const int v = m->version;
    delete m;
    OHP_FAIL(OHP_ERR_FILE_UNEXPECTED, "ohp_osh_read: unsupported format version %d", v);
  }
  if (part < 0 || part >= m->nparts) {
    const int np = m->nparts;
    delete m;
    OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_osh_read: part %d of %d", part, np);
  }
  const std::string file = d + "/" + std::to_string(part) + ".osh";
  FILE *f = fopen(file.c_str(), "rb");
  if (!f) { delete m; OHP_FAIL(OHP_ERR_FILE_OPEN, "ohp_osh_read: cannot open %s", file.c_str()); }
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf((size_t)std::max(0L, sz));
  const size_t got = fread(buf.data(), 1, buf.size(), f);
  fclose(f);
  if (got != buf.size()) { delete m; OHP_FAIL(OHP_ERR_FILE_READ, "ohp_osh_read: short read on %s", file.c_str()); }
#define BAD(...) do { ohp_set_error(__VA_ARGS__); delete m; return OHP_ERR_FILE_UNEXPECTED; } while (0)
  if (buf.size() < 2 || buf[0] != 0xa1 || buf[1] != 0x1a) BAD("ohp_osh_read: %s: bad magic", file.c_str());
  Cursor c{buf.data() + 2, buf.data() + buf.size()};
  const bool compressed = c.take<int8_t>() != 0;
  m->family = c.take<int8_t>();
  m->dim = c.take<int8_t>();
  m->comm_size = c.take<int32_t>();
  m->comm_rank = c.take<int32_t>();
  m->parting = c.take<int8_t>();
  m->nghost = c.take<int32_t>();
  const int have_hints = c.take<int8_t>();
  if (have_hints) {
    const int naxes = c.take<int32_t>();
    for (int i = 0; i < naxes * 3 && c.ok; ++i) c.take<double>();
  }
  m->nverts = c.take<int32_t>();
  if (!c.ok) BAD("ohp_osh_read: %s: truncated header", file.c_str());
  if (m->family != 0) BAD("ohp_osh_read: %s: family %d (only simplex meshes are supported)", file.c_str(), m->family);
  if (m->dim < 1 || m->dim > 3) BAD("ohp_osh_read: %s: dim %d", file.c_str(), m->dim);
  m->nents[0] = m->nverts;
  for (int dd = 1; dd <= m->dim; ++dd) {
    std::vector<uint8_t> raw;
    int64_t n = 0;
    if (!c.array(4, compressed, raw, &n)) BAD("ohp_osh_read: %s: bad down(%d) array", file.c_str(), dd);
    m->down[dd].resize((size_t)n);
    memcpy(m->down[dd].data(), raw.data(), raw.size());
    if (n % (dd + 1)) BAD("ohp_osh_read: %s: down(%d) length %lld", file.c_str(), dd, (long long)n);
    m->nents[dd] = (int32_t)(n / (dd + 1));
    if (dd > 1) {
      if (!c.array(1, compressed, raw, &n)) BAD("ohp_osh_read: %s: bad codes(%d) array", file.c_str(), dd);
      m->codes[dd].resize((size_t)n);
      memcpy(m->codes[dd].data(), raw.data(), raw.size());
    }
  }
  for (int dd = 0; dd <= m->dim; ++dd) {
    const int ntags = c.take<int32_t>();
    if (!c.ok || ntags < 0 || ntags > 4096) BAD("ohp_osh_read: %s: bad tag count at dim %d", file.c_str(), dd);
    for (int t = 0; t < ntags; ++t) {
      const int ln = c.take<int32_t>();
      if (!c.ok || ln < 0 || c.p + ln > c.end) BAD("ohp_osh_read: %s: bad tag name", file.c_str());
      std::string name((const char *)c.p, (size_t)ln);
      c.p += ln;
      Tag tag;
      tag.ncomps = c.take<int8_t>();
      tag.type = c.take<int8_t>();
      const int ts = type_size(tag.type);
      if (!c.ok || ts < 0) BAD("ohp_osh_read: %s: tag %s has unknown type %d", file.c_str(), name.c_str(), tag.type);
      if (!c.array(ts, compressed, tag.bytes, &tag.count)) BAD("ohp_osh_read: %s: bad array for tag %s", file.c_str(), name.c_str());
      m->tags[dd][name] = std::move(tag);
    }
    if (m->comm_size > 1) {
      std::vector<uint8_t> raw;
      int64_t n = 0;
      if (!c.array(4, compressed, raw, &n)) BAD("ohp_osh_read: %s: bad owner ranks", file.c_str());
      m->owner_ranks[dd].resize((size_t)n); memcpy(m->owner_ranks[dd].data(), raw.data(), raw.size());
      if (!c.array(4, compressed, raw, &n)) BAD("ohp_osh_read: %s: bad owner idxs", file.c_str());
      m->owner_idxs[dd].resize((size_t)n); memcpy(m->owner_idxs[dd].data(), raw.data(), raw.size());
    }
  }
  const int nsets = c.take<int32_t>();
  if (!c.ok) BAD("ohp_osh_read: %s: truncated trailer", file.c_str());
  if (nsets != 0) BAD("ohp_osh_read: %s: %d class sets (not supported)", file.c_str(), nsets);
  c.take<int8_t>();
  if (!c.ok || c.p != c.end) BAD("ohp_osh_read: %s: %lld trailing bytes", file.c_str(), (long long)(c.end - c.p));
  auto it = m->tags[0].find("coordinates");
  if (it == m->tags[0].end() || it->second.type != 5 || it->second.ncomps != m->dim ||
      it->second.count != (int64_t)m->nverts * m->dim)
    BAD("ohp_osh_read: %s: missing or malformed coordinates tag", file.c_str());
  for (int32_t v : m->down[1])
    if (v < 0 || v >= m->nverts) BAD("ohp_osh_read: %s: edge references vertex %d", file.c_str(), v);
#undef BAD
  const int rc = derive_elem_verts(*m);
  if (rc) { delete m; return rc; }
  *out = m;
  return 0;
}

extern "C" int ohp_osh_free(ohp_osh m) { delete m; return 0; }

extern "C" int ohp_osh_info(ohp_osh m, int *dim, int *nverts, int *nelems, int *version, int *nparts) {
  if (!m) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_osh_info: null mesh");
  if (dim) *dim = m->dim;
  if (nverts) *nverts = m->nverts;
  if (nelems) *nelems = m->nents[m->dim];
  if (version) *version = m->version;
  if (nparts) *nparts = m->nparts;
  return 0;
}
extern "C" int ohp_osh_nents(ohp_osh m, int ent_dim, int *n) {
  if (!m || ent_dim < 0 || ent_dim > m->dim) OHP_FAIL(OHP_ERR_ARG_OUTOFRANGE, "ohp_osh_nents: dim %d", ent_dim);
  *n = m->nents[ent_dim];
  return 0;
}
extern "C" int ohp_osh_elem_verts(ohp_osh m, const int32_t **ev) {
  if (!m) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_osh_elem_verts: null mesh");
  *ev = m->elem_verts.data();
  return 0;
}
extern "C" int ohp_osh_coords(ohp_osh m, const double **coords) {
  if (!m) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_osh_coords: null mesh");
  *coords = (const double *)m->tags[0]["coordinates"].bytes.data();
  return 0;
}
extern "C" int ohp_osh_tag(ohp_osh m, int ent_dim, const char *name, int *ncomps, int *type, int64_t *count,
                           const void **data) {
  if (!m || !name || ent_dim < 0 || ent_dim > m->dim) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_osh_tag: bad argument");
  auto it = m->tags[ent_dim].find(name);
  if (it == m->tags[ent_dim].end()) OHP_FAIL(OHP_ERR_ARG_WRONG, "ohp_osh_tag: no tag '%s' on dimension %d", name, ent_dim);
  if (ncomps) *ncomps = it->second.ncomps;
  if (type) *type = it->second.type;
  if (count) *count = it->second.count;
  if (data) *data = it->second.bytes.data();
  return 0;
}
