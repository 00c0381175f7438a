// ORACLE / TEST INFRASTRUCTURE ONLY.
// Stand-in for exactly the Boost.Graph surface the reference's frechet/ package touches
// (SURVEY.md Appendix A), so that /root/reference/frechet/src/{NNFrechet,LPAStar,MinPriorityQueue,
// util}.cpp compile UNMODIFIED into oracle/_ref/libref_frechet.so (Boost is absent from this image):
//   boost::adjacency_list<hash_setS, vecS, bidirectionalS, VProp, EProp>   frechet/include/util.hpp:40-42
//   graph_traits<G>::{vertex,edge}_descriptor, {vertex,edge,adjacency,in_edge}_iterator   util.hpp:43-47, util.cpp:89
//   property_map<G, T Bundle::*>::type, property_map<G, vertex_index_t>::type              util.hpp:50-64, LPAStar.h:12
//   get(&Bundle::member, g), get(vertex_index, g), map[descriptor]                         NNFrechet.cpp:100,255-256,...
//   add_vertex, add_edge (no parallel edges: hash_setS), vertices, adjacent_vertices, in_edges,
//   source, target, edge(u,v,g), vertex(i,g), tie                                          util.cpp:60-95, NNFrechet.cpp:525,649,696
// Semantics kept: vecS vertex descriptors are 0..n-1 in creation order; add_edge on a set-based
// out-edge list refuses a parallel edge and returns (existing, false); bundled properties are
// value-initialised (EProp::evaluated starts false, which NNFrechet.cpp:651-652 relies on).
// NOT kept (implementation-defined upstream): the iteration order of a hash_setS edge list, which in
// Boost depends on boost::unordered_set's bucket layout.  Here it is creation order, or reverse
// creation order with -DMINIBGL_REVERSE_ORDER (oracle/_ref/libref_frechet_rev.so).  The order only
// decides WHICH of several equally good bottleneck paths LPA* keeps; tests compare the two builds to
// show which reported quantities do not depend on it.
#ifndef ORACLE_COMPAT_BOOST_GRAPH_ADJACENCY_LIST_HPP
#define ORACLE_COMPAT_BOOST_GRAPH_ADJACENCY_LIST_HPP
// what the real header drags in and the reference relies on transitively (NNFrechet.hpp:4-8 has no
// <functional>, <random>, <unordered_map>, <memory>; NNFrechet.cpp:669 calls ::ceil)
#include <math.h>
#include <algorithm>
#include <cstddef>
#include <deque>
#include <functional>
#include <limits>
#include <memory>
#include <random>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace boost {
struct hash_setS {};
struct vecS {};
struct bidirectionalS {};
enum vertex_index_t { vertex_index };
enum edge_index_t { edge_index };

template <class OutEdgeListS, class VertexListS, class DirectedS, class VertexProperty, class EdgeProperty>
class adjacency_list {
public:
  typedef std::size_t vertex_descriptor;
  struct edge_descriptor {
    vertex_descriptor src, dst;
    std::size_t id;
    edge_descriptor() : src(0), dst(0), id((std::size_t)-1) {}
    edge_descriptor(vertex_descriptor s, vertex_descriptor d, std::size_t i) : src(s), dst(d), id(i) {}
    bool operator==(const edge_descriptor &o) const { return id == o.id; }
    bool operator!=(const edge_descriptor &o) const { return id != o.id; }
  };
  typedef VertexProperty vertex_bundled;
  typedef EdgeProperty edge_bundled;
  struct EdgeRec {
    vertex_descriptor src, dst;
    EdgeProperty prop;
  };
  struct VertexRec {
    VertexProperty prop;
    std::vector<std::size_t> out, in; // edge ids in creation order
  };
  // deque: records never move, so references handed out by the property maps stay valid
  std::deque<VertexRec> m_vertices;
  std::deque<EdgeRec> m_edges; // global creation order
};

// ---- iterators (defined in namespace boost so that the reference's unqualified tie() finds boost::tie) ----
class minibgl_counting_iterator {
public:
  minibgl_counting_iterator() : i_(0) {}
  explicit minibgl_counting_iterator(std::size_t i) : i_(i) {}
  std::size_t operator*() const { return i_; }
  minibgl_counting_iterator &operator++() {
    ++i_;
    return *this;
  }
  bool operator==(const minibgl_counting_iterator &o) const { return i_ == o.i_; }
  bool operator!=(const minibgl_counting_iterator &o) const { return i_ != o.i_; }

private:
  std::size_t i_;
};

// walks a vector of edge ids forwards (creation order) or backwards (MINIBGL_REVERSE_ORDER)
template <class G, bool kTargets> class minibgl_edge_list_iterator {
public:
  minibgl_edge_list_iterator() : g_(nullptr), ids_(nullptr), pos_(0) {}
  minibgl_edge_list_iterator(const G *g, const std::vector<std::size_t> *ids, std::size_t pos)
      : g_(g), ids_(ids), pos_(pos) {}
  typename std::conditional<kTargets, typename G::vertex_descriptor, typename G::edge_descriptor>::type
  operator*() const {
    return deref(std::integral_constant<bool, kTargets>());
  }
  minibgl_edge_list_iterator &operator++() {
    ++pos_;
    return *this;
  }
  bool operator==(const minibgl_edge_list_iterator &o) const { return pos_ == o.pos_ && ids_ == o.ids_; }
  bool operator!=(const minibgl_edge_list_iterator &o) const { return !(*this == o); }

private:
  std::size_t edge_id() const {
#ifdef MINIBGL_REVERSE_ORDER
    return (*ids_)[ids_->size() - 1 - pos_];
#else
    return (*ids_)[pos_];
#endif
  }
  typename G::vertex_descriptor deref(std::true_type) const { return g_->m_edges[edge_id()].dst; }
  typename G::edge_descriptor deref(std::false_type) const {
    const std::size_t id = edge_id();
    return typename G::edge_descriptor(g_->m_edges[id].src, g_->m_edges[id].dst, id);
  }
  const G *g_;
  const std::vector<std::size_t> *ids_;
  std::size_t pos_;
};

template <class G> struct graph_traits {
  typedef typename G::vertex_descriptor vertex_descriptor;
  typedef typename G::edge_descriptor edge_descriptor;
  typedef minibgl_counting_iterator vertex_iterator;
  typedef minibgl_edge_list_iterator<G, false> edge_iterator; // named by util.hpp:46, never used
  typedef minibgl_edge_list_iterator<G, true> adjacency_iterator;
  typedef minibgl_edge_list_iterator<G, false> in_edge_iterator;
  typedef minibgl_edge_list_iterator<G, false> out_edge_iterator;
};

// ---- property maps ----
template <class G, class Bundle, class T> class minibgl_bundle_map {
public:
  minibgl_bundle_map() : g_(nullptr), pm_(nullptr) {}
  minibgl_bundle_map(G *g, T Bundle::*pm) : g_(g), pm_(pm) {}
  // vertex bundle
  template <class B = Bundle>
  typename std::enable_if<std::is_same<B, typename G::vertex_bundled>::value, T &>::type
  operator[](typename G::vertex_descriptor v) const {
    return g_->m_vertices[v].prop.*pm_;
  }
  // edge bundle
  template <class B = Bundle>
  typename std::enable_if<std::is_same<B, typename G::edge_bundled>::value, T &>::type
  operator[](const typename G::edge_descriptor &e) const {
    return g_->m_edges[e.id].prop.*pm_;
  }

private:
  G *g_;
  T Bundle::*pm_;
};
struct minibgl_identity_map {
  std::size_t operator[](std::size_t v) const { return v; }
};

template <class G, class Tag> struct property_map;
template <class G, class T, class Bundle> struct property_map<G, T Bundle::*> {
  typedef minibgl_bundle_map<G, Bundle, T> type;
  typedef minibgl_bundle_map<const G, Bundle, const T> const_type;
};
template <class G> struct property_map<G, vertex_index_t> {
  typedef minibgl_identity_map type;
  typedef minibgl_identity_map const_type;
};

#define MINIBGL_TPL template <class O, class VL, class D, class VP, class EP>
#define MINIBGL_G adjacency_list<O, VL, D, VP, EP>

template <class T, class Bundle, class O, class VL, class D, class VP, class EP>
minibgl_bundle_map<MINIBGL_G, Bundle, T> get(T Bundle::*pm, MINIBGL_G &g) {
  return minibgl_bundle_map<MINIBGL_G, Bundle, T>(&g, pm);
}
MINIBGL_TPL minibgl_identity_map get(vertex_index_t, const MINIBGL_G &) { return minibgl_identity_map(); }

// ---- mutators / queries ----
MINIBGL_TPL typename MINIBGL_G::vertex_descriptor add_vertex(MINIBGL_G &g) {
  g.m_vertices.emplace_back(); // value-initialises the bundle
  g.m_vertices.back().prop = VP();
  return g.m_vertices.size() - 1;
}

MINIBGL_TPL std::pair<typename MINIBGL_G::edge_descriptor, bool>
edge(typename MINIBGL_G::vertex_descriptor u, typename MINIBGL_G::vertex_descriptor v, const MINIBGL_G &g) {
  typedef typename MINIBGL_G::edge_descriptor E;
  for (std::size_t id : g.m_vertices[u].out)
    if (g.m_edges[id].dst == v)
      return std::make_pair(E(u, v, id), true);
  return std::make_pair(E(), false);
}

MINIBGL_TPL std::pair<typename MINIBGL_G::edge_descriptor, bool>
add_edge(typename MINIBGL_G::vertex_descriptor u, typename MINIBGL_G::vertex_descriptor v, MINIBGL_G &g) {
  typedef typename MINIBGL_G::edge_descriptor E;
  std::pair<E, bool> found = edge(u, v, g);
  if (found.second) // set-based out-edge list: no parallel edges
    return std::make_pair(found.first, false);
  typename MINIBGL_G::EdgeRec rec;
  rec.src = u;
  rec.dst = v;
  rec.prop = EP(); // value-initialised: EProp::evaluated == false
  g.m_edges.push_back(rec);
  const std::size_t id = g.m_edges.size() - 1;
  g.m_vertices[u].out.push_back(id);
  g.m_vertices[v].in.push_back(id);
  return std::make_pair(E(u, v, id), true);
}

MINIBGL_TPL std::pair<minibgl_counting_iterator, minibgl_counting_iterator> vertices(const MINIBGL_G &g) {
  return std::make_pair(minibgl_counting_iterator(0), minibgl_counting_iterator(g.m_vertices.size()));
}
MINIBGL_TPL std::pair<minibgl_edge_list_iterator<MINIBGL_G, true>, minibgl_edge_list_iterator<MINIBGL_G, true>>
adjacent_vertices(typename MINIBGL_G::vertex_descriptor v, const MINIBGL_G &g) {
  typedef minibgl_edge_list_iterator<MINIBGL_G, true> It;
  const std::vector<std::size_t> *l = &g.m_vertices[v].out;
  return std::make_pair(It(&g, l, 0), It(&g, l, l->size()));
}
MINIBGL_TPL std::pair<minibgl_edge_list_iterator<MINIBGL_G, false>, minibgl_edge_list_iterator<MINIBGL_G, false>>
in_edges(typename MINIBGL_G::vertex_descriptor v, const MINIBGL_G &g) {
  typedef minibgl_edge_list_iterator<MINIBGL_G, false> It;
  const std::vector<std::size_t> *l = &g.m_vertices[v].in;
  return std::make_pair(It(&g, l, 0), It(&g, l, l->size()));
}
MINIBGL_TPL typename MINIBGL_G::vertex_descriptor source(const typename MINIBGL_G::edge_descriptor &e,
                                                        const MINIBGL_G &) {
  return e.src;
}
MINIBGL_TPL typename MINIBGL_G::vertex_descriptor target(const typename MINIBGL_G::edge_descriptor &e,
                                                        const MINIBGL_G &) {
  return e.dst;
}
MINIBGL_TPL typename MINIBGL_G::vertex_descriptor vertex(std::size_t i, const MINIBGL_G &) { return i; }
MINIBGL_TPL std::size_t num_vertices(const MINIBGL_G &g) { return g.m_vertices.size(); }
MINIBGL_TPL std::size_t num_edges(const MINIBGL_G &g) { return g.m_edges.size(); }
#undef MINIBGL_TPL
#undef MINIBGL_G

// ---- tie(a, b) = pair ----
template <class A, class B> struct minibgl_tier {
  A &a;
  B &b;
  minibgl_tier(A &a_, B &b_) : a(a_), b(b_) {}
  template <class U, class V> minibgl_tier &operator=(const std::pair<U, V> &p) {
    a = p.first;
    b = p.second;
    return *this;
  }
};
template <class A, class B> minibgl_tier<A, B> tie(A &a, B &b) { return minibgl_tier<A, B>(a, b); }
} // namespace boost
#endif
