// TEST INFRASTRUCTURE ONLY -- stand-in for nanoflann's KDTreeVectorOfVectorsAdaptor.h (un-vendored by the reference,
// .gitignore:76), written from nanoflann's published interface so that FJSPH's OWN sources can be compiled where they
// lie under /root/reference/src into oracle/_ref/ (oracle/Makefile.ref).  Not nanoflann: no KD-tree, an exhaustive
// scan with the same results:
//   radiusSearch  : every point with metric_L2_Simple distance  d2 = sum_d (q_d - p_d)^2  (accumulated d = 0,1,2)
//                   STRICTLY below the radius argument, (index, d2) pairs.  Returned in ascending index order (nanoflann
//                   with SearchParameters(0, false) returns tree-traversal order; the reference never relies on it).
//   findNeighbors : KNNResultSet semantics, the k smallest (d2, index) in ascending order, ties by lower index.
// Only the members the reference touches exist (Var.h:889-899, Neighbours.cpp:10,44,60, Containment.cpp:441-444,
// FJSPH.cpp:145-149).
#pragma once
#include <algorithm>
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
        void buildIndex() {} /* the scan reads the live container */

        template <class Alloc>
        size_t radiusSearch(const num_t* query, const num_t& radius, std::vector<nanoflann::ResultItem<IndexType, num_t>, Alloc>& out,
                            const nanoflann::SearchParameters& = nanoflann::SearchParameters()) const
        {
            const VectorOfVectorsType& pts = owner.m_data;
            const size_t n = pts.size();
            out.clear();
            for (size_t j = 0; j < n; ++j)
            {
                num_t d2 = num_t(0);
                for (int d = 0; d < DIM; ++d)
                {
                    const num_t diff = query[d] - pts[j][d];
                    d2 += diff * diff;
                }
                if (d2 < radius)
                    out.emplace_back(IndexType(j), d2);
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
                num_t d2 = num_t(0);
                for (int d = 0; d < DIM; ++d)
                {
                    const num_t diff = query[d] - pts[j][d];
                    d2 += diff * diff;
                }
                if (d2 < result.worstDist() || !result.full())
                    result.addPoint(d2, IndexType(j));
            }
            return result.full();
        }
        const self_t& owner;
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
