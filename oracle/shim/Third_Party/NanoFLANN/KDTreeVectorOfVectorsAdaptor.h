// TEST INFRASTRUCTURE ONLY -- stand-in for nanoflann's KDTreeVectorOfVectorsAdaptor.h (un-vendored by the reference,
// .gitignore:76), written from nanoflann's published interface so that FJSPH's OWN sources can be compiled where they
// lie under /root/reference/src into oracle/_ref/ (oracle/Makefile.ref).  Not nanoflann: no KD-tree, an exhaustive
// scan with the same results:
//   radiusSearch  : every point with metric_L2_Simple distance  d2 = sum_d (q_d - p_d)^2  (accumulated d = 0,1,2)
//                   STRICTLY below the radius argument, (index, d2) pairs.  Returned in grid-cell order, ascending index
//                   inside a cell (nanoflann with SearchParameters(0, false) returns tree-traversal order; the reference
//                   never relies on it).
//   findNeighbors : KNNResultSet semantics, the k smallest (d2, index) in ascending order, ties by lower index.
// Only the members the reference touches exist (Var.h:889-899, Neighbours.cpp:10,44,60, Containment.cpp:441-444,
// FJSPH.cpp:145-149).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <limits>
#include <utility>
#include <vector>

namespace nanoflann
{
struct metric_L2_Simple
{
};

template <typename IndexType = size_t, typename DistanceType = double>
struct ResultItem
{
    ResultItem() = default;
    ResultItem(const IndexType index, const DistanceType distance) : first(index), second(distance) {}
    IndexType first;
    DistanceType second;
};

struct SearchParameters
{
    SearchParameters(float eps_ = 0, bool sorted_ = true) : eps(eps_), sorted(sorted_) {}
    float eps;
    bool sorted;
};

struct KDTreeSingleIndexAdaptorParams
{
    KDTreeSingleIndexAdaptorParams(size_t leaf = 10) : leaf_max_size(leaf) {}
    size_t leaf_max_size;
};

template <typename DistanceType, typename IndexType = size_t, typename CountType = size_t>
class KNNResultSet
{
  public:
    explicit KNNResultSet(CountType capacity_) : indices(nullptr), dists(nullptr), capacity(capacity_), count(0) {}
    void init(IndexType* indices_, DistanceType* dists_)
    {
        indices = indices_;
        dists = dists_;
        count = 0;
        if (capacity)
            dists[capacity - 1] = (std::numeric_limits<DistanceType>::max)();
    }
    CountType size() const { return count; }
    bool full() const { return count == capacity; }
    bool addPoint(DistanceType dist, IndexType index)
    {
        CountType i;
        for (i = count; i > 0; --i)
        {
            if ((dists[i - 1] > dist) || ((dist == dists[i - 1]) && (indices[i - 1] > index)))
            {
                if (i < capacity)
                {
                    dists[i] = dists[i - 1];
                    indices[i] = indices[i - 1];
                }
            }
            else
                break;
        }
        if (i < capacity)
        {
            dists[i] = dist;
            indices[i] = index;
        }
        if (count < capacity)
            count++;
        return true;
    }
    DistanceType worstDist() const { return dists[capacity - 1]; }

  private:
    IndexType* indices;
    DistanceType* dists;
    CountType capacity;
    CountType count;
};
} // namespace nanoflann

template <class VectorOfVectorsType, typename num_t = double, int DIM = -1, class Distance = nanoflann::metric_L2_Simple,
          typename IndexType = size_t>
struct KDTreeVectorOfVectorsAdaptor
{
    using self_t = KDTreeVectorOfVectorsAdaptor<VectorOfVectorsType, num_t, DIM, Distance, IndexType>;

    struct index_t
    {
        explicit index_t(const self_t& owner_) : owner(owner_) {}

        /* buildIndex bins the points of the live container into a uniform grid (cell edge chosen so that a cell holds a
           few points); radiusSearch visits the cells the ball touches, in ascending cell order, and reports each cell's
           points in ascending index order.  A search made after the container changed without a buildIndex (the
           reference never does that) falls back to the exhaustive scan. */
        void buildIndex()
        {
            const VectorOfVectorsType& pts = owner.m_data;
            const size_t n = pts.size();
            built_n = n;
            start.clear();
            order.clear();
            if (n < 64)
                return;
            for (int d = 0; d < DIM; ++d)
            {
                lo[d] = hi[d] = pts[0][d];
            }
            for (size_t j = 1; j < n; ++j)
                for (int d = 0; d < DIM; ++d)
                {
                    const num_t x = pts[j][d];
                    lo[d] = x < lo[d] ? x : lo[d];
                    hi[d] = x > hi[d] ? x : hi[d];
                }
            num_t vol = num_t(1);
            for (int d = 0; d < DIM; ++d) vol *= (hi[d] - lo[d]) > num_t(0) ? (hi[d] - lo[d]) : num_t(1);
            edge = std::pow(vol * num_t(8) / num_t(n), num_t(1) / num_t(DIM)); /* ~8 points per cell */
            if (!(edge > num_t(0)))
            {
                built_n = 0;
                return;
            }
            size_t total = 1;
            for (int d = 0; d < DIM; ++d)
            {
                dims[d] = size_t((hi[d] - lo[d]) / edge) + 1;
                total *= dims[d];
            }
            if (total > 8 * n + 1024)
            {
                built_n = 0; /* very elongated clouds: keep the scan */
                return;
            }
            std::vector<size_t> cell(n);
            start.assign(total + 1, 0);
            for (size_t j = 0; j < n; ++j)
            {
                cell[j] = cell_of(&pts[j]);
                start[cell[j] + 1]++;
            }
            for (size_t c = 0; c < total; ++c) start[c + 1] += start[c];
            order.resize(n);
            std::vector<size_t> fill(start.begin(), start.end() - 1);
            for (size_t j = 0; j < n; ++j) order[fill[cell[j]]++] = j; /* ascending j inside a cell */
        }

        template <class P>
        size_t cell_of(const P* p) const
        {
            size_t c = 0, stride = 1;
            for (int d = 0; d < DIM; ++d)
            {
                size_t k = size_t(((*p)[d] - lo[d]) / edge);
                k = k < dims[d] ? k : dims[d] - 1;
                c += k * stride;
                stride *= dims[d];
            }
            return c;
        }

        static num_t dist2(const num_t* query, const VectorOfVectorsType& pts, size_t j)
        {
            num_t d2 = num_t(0);
            for (int d = 0; d < DIM; ++d)
            {
                const num_t diff = query[d] - pts[j][d];
                d2 += diff * diff;
            }
            return d2;
        }

        template <class Alloc>
        size_t radiusSearch(const num_t* query, const num_t& radius, std::vector<nanoflann::ResultItem<IndexType, num_t>, Alloc>& out,
                            const nanoflann::SearchParameters& = nanoflann::SearchParameters()) const
        {
            const VectorOfVectorsType& pts = owner.m_data;
            const size_t n = pts.size();
            out.clear();
            if (order.empty() || built_n != n)
            {
                for (size_t j = 0; j < n; ++j)
                {
                    const num_t d2 = dist2(query, pts, j);
                    if (d2 < radius)
                        out.emplace_back(IndexType(j), d2);
                }
                return out.size();
            }
            /* radius is the SQUARED search radius (metric_L2_Simple) */
            const num_t r = std::sqrt(radius) * (num_t(1) + num_t(1e-12));
            long k0[3] = {0, 0, 0}, k1[3] = {0, 0, 0};
            for (int d = 0; d < DIM; ++d)
            {
                long a = long(std::floor((query[d] - r - lo[d]) / edge)), b = long(std::floor((query[d] + r - lo[d]) / edge));
                a = a < 0 ? 0 : a;
                b = b >= long(dims[d]) ? long(dims[d]) - 1 : b;
                k0[d] = a;
                k1[d] = b;
            }
            for (long kz = k0[2]; kz <= k1[2]; ++kz)
                for (long ky = k0[1]; ky <= k1[1]; ++ky)
                    for (long kx = k0[0]; kx <= k1[0]; ++kx)
                    {
                        size_t c = size_t(kx);
                        if (DIM > 1)
                            c += size_t(ky) * dims[0];
                        if (DIM > 2)
                            c += size_t(kz) * dims[0] * dims[1 % DIM];
                        for (size_t s = start[c]; s < start[c + 1]; ++s)
                        {
                            const size_t j = order[s];
                            const num_t d2 = dist2(query, pts, j);
                            if (d2 < radius)
                                out.emplace_back(IndexType(j), d2);
                        }
                    }
            return out.size();
        }

        template <class RESULTSET>
        bool findNeighbors(RESULTSET& result, const num_t* query,
                           const nanoflann::SearchParameters& = nanoflann::SearchParameters()) const
        {
            const VectorOfVectorsType& pts = owner.m_data;
            const size_t n = pts.size();
            for (size_t j = 0; j < n; ++j)
            {
                const num_t d2 = dist2(query, pts, j);
                if (d2 < result.worstDist() || !result.full())
                    result.addPoint(d2, IndexType(j));
            }
            return result.full();
        }
        const self_t& owner;
        size_t built_n = 0;
        num_t lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, edge = 0;
        size_t dims[3] = {1, 1, 1};
        std::vector<size_t> start, order;
    };

    KDTreeVectorOfVectorsAdaptor(const size_t /*dimensionality*/, const VectorOfVectorsType& mat, const int /*leaf_max_size*/ = 10,
                                 const unsigned int /*n_thread_build*/ = 1)
        : m_data(mat)
    {
        index = new index_t(*this);
    }
    ~KDTreeVectorOfVectorsAdaptor() { delete index; }
    KDTreeVectorOfVectorsAdaptor(const self_t&) = delete;

    index_t* index;
    const VectorOfVectorsType& m_data;
};
