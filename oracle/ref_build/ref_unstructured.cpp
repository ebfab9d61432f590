// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  C ABI over the reference's vendored toy mesh
// library and its pre-generated / hand-written naive-ico stencils, all compiled UNMODIFIED from
// /root/reference:
//   dawn/src/toylib/toylib.{hpp,cpp}                       (mesh + field containers)
//   dawn/src/interface/toylib_interface.hpp                (getNeighbors / reduce, :95-323)
//   dawn/test/integration-test/unstructured/reference/reference_{diffusion,gradient,diamond,
//        diamondWeights,intp}.hpp
//   dawn/test/integration-test/dawn4py-tests/data/ICON_laplacian_stencil_ref.cpp  (nabla2)
//   dawn/test/integration-test/dawn4py-tests/data/generate_versioned_field_ref.cpp (versioned field,
//        nested conditionals on field values)
//
// Plain-buffer conventions: dense fields are level-major [k][n] with n = toylib element id
// (edges: id in the all_edges() slot numbering, invalid slots simply unused); sparse fields are
// [k][n][sparse].
#include <cassert>
#include <chrono>
#include <cstring>
#include <vector>

#define DAWN_GENERATED 1
#include <interface/toylib_interface.hpp>

// ICON_laplacian_stencil_ref.cpp:168-171 needs allocateField()/numEdges() from the mesh tag; the
// toylib adapter has numEdges but no allocateField, so provide one (sized by the mesh the wrapper
// is currently running on).
namespace toylibInterface {
static const toylib::Grid* g_current_grid = nullptr;
inline toylib::EdgeData<double> allocateField(toylibTag, std::size_t, int k_size) {
  return toylib::EdgeData<double>(*g_current_grid, k_size);
}
} // namespace toylibInterface

#include REF_UNSTR(reference_diffusion.hpp)
#include REF_UNSTR(reference_gradient.hpp)
#include REF_UNSTR(reference_diamond.hpp)
#include REF_UNSTR(reference_diamondWeights.hpp)
#include REF_UNSTR(reference_intp.hpp)
#include REF_ICONLAP
#include REF_VERSIONED

namespace {
using dawn::LocationType;
using toylibInterface::toylibTag;

toylib::Grid make_grid(int nx, int ny, double lx, double ly, int equilat) {
  return toylib::Grid(nx, ny, false, lx, ly, equilat != 0);
}

template <class Data>
void load(Data& d, const double* src, int n, int k_size) {
  int k = 0;
  for(auto& level : d) {
    std::memcpy(level.data(), src + std::size_t(k) * n, sizeof(double) * n);
    ++k;
  }
  (void)k_size;
}
template <class Data>
void store(Data& d, double* dst, int n) {
  int k = 0;
  for(auto& level : d) {
    std::memcpy(dst + std::size_t(k) * n, level.data(), sizeof(double) * n);
    ++k;
  }
}
} // namespace

extern "C" {

int ref_toylib_sizes(int nx, int ny, int* n_vertices, int* n_edge_slots, int* n_cells,
                     int* n_valid_edges) {
  auto g = make_grid(nx, ny, 0, 0, 0);
  *n_vertices = (int)g.vertices().size();
  *n_edge_slots = (int)g.all_edges().size();
  *n_cells = (int)g.faces().size();
  *n_valid_edges = (int)g.edges().size();
  return 0;
}

// ids of valid edges in iteration order of getEdges(toylibTag, mesh)
int ref_toylib_valid_edges(int nx, int ny, int* ids) {
  auto g = make_grid(nx, ny, 0, 0, 0);
  int n = 0;
  for(auto e : getEdges(toylibTag{}, g))
    ids[n++] = e->id();
  return n;
}

int ref_toylib_vertex_xy(int nx, int ny, double lx, double ly, int equilat, double* x, double* y) {
  auto g = make_grid(nx, ny, lx, ly, equilat);
  for(auto const& v : g.vertices()) {
    x[v.id()] = v.x();
    y[v.id()] = v.y();
  }
  return 0;
}

// Neighbour table for `chain` (LocationType ints: 0 cells, 1 edges, 2 vertices), row-major
// table[elem_id * max_nbh + n], padded with -1, in the order getNeighbors() returns them
// (toylib_interface.hpp:144-274).  Rows of invalid edge slots stay -1.  Returns the largest
// neighbour count seen, or -1 if it exceeds max_nbh.
int ref_toylib_nbh_table(int nx, int ny, const int* chain, int chain_len, int max_nbh, int* table) {
  auto g = make_grid(nx, ny, 0, 0, 0);
  std::vector<LocationType> ch;
  for(int c = 0; c < chain_len; ++c)
    ch.push_back(static_cast<LocationType>(chain[c]));
  std::vector<const toylib::ToylibElement*> elems;
  std::size_t n_rows = 0;
  switch(ch.front()) {
  case LocationType::Cells:
    elems = getCells(toylibTag{}, g);
    n_rows = g.faces().size();
    break;
  case LocationType::Edges:
    elems = getEdges(toylibTag{}, g);
    n_rows = g.all_edges().size();
    break;
  case LocationType::Vertices:
    elems = getVertices(toylibTag{}, g);
    n_rows = g.vertices().size();
    break;
  }
  for(std::size_t i = 0; i < n_rows * max_nbh; ++i)
    table[i] = -1;
  int seen = 0;
  for(auto e : elems) {
    auto nbh = getNeighbors(toylibTag{}, g, ch, e);
    if((int)nbh.size() > max_nbh)
      return -1;
    seen = std::max(seen, (int)nbh.size());
    for(std::size_t n = 0; n < nbh.size(); ++n)
      table[std::size_t(e->id()) * max_nbh + n] = nbh[n]->id();
  }
  return seen;
}

int ref_unstr_diffusion(int nx, int ny, int k_size, const double* in, double* out) {
  auto g = make_grid(nx, ny, 0, 0, 0);
  int nc = (int)g.faces().size();
  toylib::FaceData<double> fin(g, k_size), fout(g, k_size);
  load(fin, in, nc, k_size);
  load(fout, out, nc, k_size);
  dawn_generated::cxxnaiveico::reference_diffusion<toylibTag> st(g, k_size, fin, fout);
  st.run();
  store(fout, out, nc);
  return 0;
}

int ref_unstr_gradient(int nx, int ny, int k_size, double* cell_field, double* edge_field) {
  auto g = make_grid(nx, ny, 0, 0, 0);
  int nc = (int)g.faces().size(), ne = (int)g.all_edges().size();
  toylib::FaceData<double> fc(g, k_size);
  toylib::EdgeData<double> fe(g, k_size);
  load(fc, cell_field, nc, k_size);
  load(fe, edge_field, ne, k_size);
  dawn_generated::cxxnaiveico::reference_gradient<toylibTag> st(g, k_size, fc, fe);
  st.run();
  store(fc, cell_field, nc);
  store(fe, edge_field, ne);
  return 0;
}

int ref_unstr_diamond(int nx, int ny, int k_size, double* edge_field, const double* vertex_field) {
  auto g = make_grid(nx, ny, 0, 0, 0);
  int nv = (int)g.vertices().size(), ne = (int)g.all_edges().size();
  toylib::EdgeData<double> fe(g, k_size);
  toylib::VertexData<double> fv(g, k_size);
  load(fe, edge_field, ne, k_size);
  load(fv, vertex_field, nv, k_size);
  dawn_generated::cxxnaiveico::reference_diamond<toylibTag> st(g, k_size, fe, fv);
  st.run();
  store(fe, edge_field, ne);
  return 0;
}

int ref_unstr_diamond_weights(int nx, int ny, int k_size, double* out, const double* inv_edge_length,
                              const double* inv_vert_length, const double* in) {
  auto g = make_grid(nx, ny, 0, 0, 0);
  int nv = (int)g.vertices().size(), ne = (int)g.all_edges().size();
  toylib::EdgeData<double> fo(g, k_size), fel(g, k_size), fvl(g, k_size);
  toylib::VertexData<double> fin(g, k_size);
  load(fo, out, ne, k_size);
  load(fel, inv_edge_length, ne, k_size);
  load(fvl, inv_vert_length, ne, k_size);
  load(fin, in, nv, k_size);
  dawn_generated::cxxnaiveico::reference_diamondWeights<toylibTag> st(g, k_size, fo, fel, fvl, fin);
  st.run();
  store(fo, out, ne);
  return 0;
}

int ref_unstr_intp(int nx, int ny, int k_size, const double* in, double* out) {
  auto g = make_grid(nx, ny, 0, 0, 0);
  int nc = (int)g.faces().size();
  toylib::FaceData<double> fin(g, k_size), fout(g, k_size);
  load(fin, in, nc, k_size);
  load(fout, out, nc, k_size);
  dawn_generated::cxxnaiveico::reference_intp<toylibTag> st(g, k_size, fin, fout);
  st.run();
  store(fout, out, nc);
  return 0;
}

// nabla2: API order of ICON_laplacian_stencil_ref.cpp:155-164.
int ref_icon_laplacian(int nx, int ny, int k_size, const double* vec, double* div_vec,
                       double* rot_vec, double* nabla2_vec, const double* primal_edge_length,
                       const double* dual_edge_length, const double* tangent_orientation,
                       const double* geofac_rot /*[k][V][6]*/,
                       const double* geofac_div /*[k][C][3]*/) {
  auto g = make_grid(nx, ny, 0, 0, 0);
  toylibInterface::g_current_grid = &g;
  int nv = (int)g.vertices().size(), ne = (int)g.all_edges().size(), nc = (int)g.faces().size();
  toylib::EdgeData<double> fvec(g, k_size), fnab(g, k_size), fpel(g, k_size), fdel(g, k_size),
      ftan(g, k_size);
  toylib::FaceData<double> fdiv(g, k_size);
  toylib::VertexData<double> frot(g, k_size);
  toylib::SparseVertexData<double> grot(g, 6, k_size);
  toylib::SparseFaceData<double> gdiv(g, 3, k_size);
  load(fvec, vec, ne, k_size);
  load(fnab, nabla2_vec, ne, k_size);
  load(fpel, primal_edge_length, ne, k_size);
  load(fdel, dual_edge_length, ne, k_size);
  load(ftan, tangent_orientation, ne, k_size);
  load(fdiv, div_vec, nc, k_size);
  load(frot, rot_vec, nv, k_size);
  for(int k = 0; k < k_size; ++k) {
    for(auto const& v : g.vertices())
      for(int s = 0; s < 6; ++s)
        grot(v, s, k) = geofac_rot[(std::size_t(k) * nv + v.id()) * 6 + s];
    for(auto const& f : g.faces())
      for(int s = 0; s < 3; ++s)
        gdiv(f, s, k) = geofac_div[(std::size_t(k) * nc + f.id()) * 3 + s];
  }
  dawn_generated::cxxnaiveico::ICON_laplacian_stencil<toylibTag> st(
      g, k_size, fvec, fdiv, frot, fnab, fpel, fdel, ftan, grot, gdiv);
  st.run();
  store(fdiv, div_vec, nc);
  store(frot, rot_vec, nv);
  store(fnab, nabla2_vec, ne);
  toylibInterface::g_current_grid = nullptr;
  return 0;
}

// generate_versioned_field_ref.cpp:103-111: five edge fields, the versioned copy c_0 is allocated inside.
int ref_generate_versioned_field(int nx, int ny, int k_size, double* a, const double* b, double* c,
                                 const double* d, const double* e) {
  auto g = make_grid(nx, ny, 0, 0, 0);
  toylibInterface::g_current_grid = &g;
  int ne = (int)g.all_edges().size();
  toylib::EdgeData<double> fa(g, k_size), fb(g, k_size), fc(g, k_size), fd(g, k_size), fe(g, k_size);
  load(fa, a, ne, k_size);
  load(fb, b, ne, k_size);
  load(fc, c, ne, k_size);
  load(fd, d, ne, k_size);
  load(fe, e, ne, k_size);
  dawn_generated::cxxnaiveico::generate_versioned_field<toylibTag> st(g, k_size, fa, fb, fc, fd, fe);
  st.run();
  store(fa, a, ne);
  store(fc, c, ne);
  toylibInterface::g_current_grid = nullptr;
  return 0;
}

// CPU baseline of bench.py's ICON workload: the reference's generated naive-ico nabla2 exactly as it
// ships (single thread, toylib mesh), `reps` calls of run() timed on their own - mesh construction and
// field set-up stay outside.  Returns the number of edges (negative on failure); seconds[r] = wall
// time of call r.  Edge count = valid edges, the unit of bench.py's metric.
int ref_icon_laplacian_timed(int nx, int ny, int k_size, int reps, double* seconds) {
  auto g = make_grid(nx, ny, 0, 0, 0);
  toylibInterface::g_current_grid = &g;
  toylib::EdgeData<double> fvec(g, k_size), fnab(g, k_size), fpel(g, k_size), fdel(g, k_size),
      ftan(g, k_size);
  toylib::FaceData<double> fdiv(g, k_size);
  toylib::VertexData<double> frot(g, k_size);
  toylib::SparseVertexData<double> grot(g, 6, k_size);
  toylib::SparseFaceData<double> gdiv(g, 3, k_size);
  for(int k = 0; k < k_size; ++k) {
    for(toylib::Edge const& e : g.edges()) {
      fvec(e, k) = 1.0 + 1e-3 * (e.id() % 97) + 1e-2 * k;
      fpel(e, k) = 1.0 + 1e-3 * (e.id() % 89);
      fdel(e, k) = 1.0 + 1e-3 * (e.id() % 83);
      ftan(e, k) = (e.id() & 1) ? 1.0 : -1.0;
    }
    for(auto const& v : g.vertices())
      for(int s = 0; s < 6; ++s)
        grot(v, s, k) = 0.1 * (s + 1);
    for(auto const& f : g.faces())
      for(int s = 0; s < 3; ++s)
        gdiv(f, s, k) = 0.2 * (s + 1);
  }
  dawn_generated::cxxnaiveico::ICON_laplacian_stencil<toylibTag> st(
      g, k_size, fvec, fdiv, frot, fnab, fpel, fdel, ftan, grot, gdiv);
  for(int r = 0; r < reps; ++r) {
    auto t0 = std::chrono::steady_clock::now();
    st.run();
    seconds[r] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  toylibInterface::g_current_grid = nullptr;
  return (int)g.edges().size();
}
}
