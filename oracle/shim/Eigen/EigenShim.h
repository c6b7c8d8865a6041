// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for the parts of Eigen 3.4.0 that ArtNlk/FlipSolver2d's
// sources name (FlipSolver2dLib/CMakeLists.txt:100-111 fetches Eigen 3.4.0 with
// FetchContent; there is no network here, so the real headers are absent).
// It exists so that the UNMODIFIED reference sources under /root/reference can be
// compiled into oracle/_ref/ as the parity oracle and CPU baseline.
//
// What is real here:
//   * VectorXd / VectorXi              (resize, size, [], (), Constant)
//   * SparseMatrix<S,Opt,SI>           (resize, reserve, coeffRef, makeCompressed ...)
//   * ConjugateGradient<M, UpLo>       Eigen 3.4's conjugate_gradient() with the
//                                      DiagonalPreconditioner on a self-adjoint view
//                                      (x0 = 0, threshold tol^2*|b|^2, maxIter = 2n).
//                                      This is a restatement of Eigen's published
//                                      algorithm -> viscosity parity is "unpinned".
// Everything the never-instantiated InversePoissonPreconditioner template
// (FlipSolver2dLib/inversepoissonpreconditioner.h) touches only has to parse.
#ifndef FS2D_ORACLE_EIGEN_SHIM_H
#define FS2D_ORACLE_EIGEN_SHIM_H

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <limits>
#include <utility>
#include <vector>

#define EIGEN_CONSTEXPR constexpr
#define EIGEN_NOEXCEPT noexcept
#define eigen_assert(x) assert(x)

namespace Eigen
{
typedef std::ptrdiff_t Index;
const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1 };
enum { Lower = 1, Upper = 2, StrictlyLower = 9, StrictlyUpper = 10 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };

inline void initParallel() {}
inline void setNbThreads(int) {}

template <typename T> struct NumTraits { typedef T Real; };

template <typename I> struct AMDOrdering
{
    struct PermutationType { typedef I StorageIndex; };
};

template <typename Derived> class SparseSolverBase
{
protected:
    mutable bool m_isInitialized = false;
};

template <typename S> class DenseVector
{
public:
    typedef S Scalar;
    DenseVector() = default;
    explicit DenseVector(Index n) : m_v(static_cast<size_t>(n)) {}
    void resize(Index n) { m_v.resize(static_cast<size_t>(n)); }
    Index size() const { return static_cast<Index>(m_v.size()); }
    S &operator[](Index i) { return m_v[static_cast<size_t>(i)]; }
    const S &operator[](Index i) const { return m_v[static_cast<size_t>(i)]; }
    S &operator()(Index i) { return m_v[static_cast<size_t>(i)]; }
    const S &operator()(Index i) const { return m_v[static_cast<size_t>(i)]; }
    static DenseVector Constant(Index n, S value)
    {
        DenseVector out(n);
        std::fill(out.m_v.begin(), out.m_v.end(), value);
        return out;
    }
    void setZero() { std::fill(m_v.begin(), m_v.end(), S(0)); }
    S dot(const DenseVector &o) const
    {
        S acc = S(0);
        for (size_t i = 0; i < m_v.size(); i++) acc += m_v[i] * o.m_v[i];
        return acc;
    }
    S squaredNorm() const { return dot(*this); }
    std::vector<S> &raw() { return m_v; }
    const std::vector<S> &raw() const { return m_v; }

private:
    std::vector<S> m_v;
};

typedef DenseVector<double> VectorXd;
typedef DenseVector<int> VectorXi;

// Row-list sparse matrix. Storage order is irrelevant for the semantics the
// reference relies on ((row, col) addressed coeffRef + products), so both
// RowMajor and ColMajor instantiations share this representation.
template <typename S, int Options = ColMajor, typename SI = int> class SparseMatrix
{
public:
    typedef S Scalar;
    typedef SI StorageIndex;
    typedef std::pair<Index, S> Entry;

    SparseMatrix() = default;
    SparseMatrix(Index r, Index c) { resize(r, c); }
    template <int O2, typename SI2> SparseMatrix(const SparseMatrix<S, O2, SI2> &o) { copyFrom(o); }
    template <int O2, typename SI2> SparseMatrix &operator=(const SparseMatrix<S, O2, SI2> &o)
    {
        copyFrom(o);
        return *this;
    }

    void resize(Index r, Index c)
    {
        m_rows = r;
        m_cols = c;
        m_data.assign(static_cast<size_t>(r), std::vector<Entry>());
    }
    template <typename V> void reserve(const V &perRow)
    {
        for (Index r = 0; r < m_rows && r < perRow.size(); r++)
            m_data[static_cast<size_t>(r)].reserve(static_cast<size_t>(perRow[r]));
    }
    void makeCompressed() {}
    Index rows() const { return m_rows; }
    Index cols() const { return m_cols; }
    Index outerSize() const { return m_rows; }

    S &coeffRef(Index r, Index c)
    {
        std::vector<Entry> &row = m_data[static_cast<size_t>(r)];
        auto it = std::lower_bound(row.begin(), row.end(), c,
                                   [](const Entry &e, Index col) { return e.first < col; });
        if (it == row.end() || it->first != c) it = row.insert(it, Entry(c, S(0)));
        return it->second;
    }

    void setIdentity()
    {
        for (Index r = 0; r < m_rows; r++)
        {
            m_data[static_cast<size_t>(r)].clear();
            if (r < m_cols) m_data[static_cast<size_t>(r)].push_back(Entry(r, S(1)));
        }
    }

    class InnerIterator
    {
    public:
        InnerIterator(SparseMatrix &m, Index outer) : m_row(&m.m_data[static_cast<size_t>(outer)]), m_outer(outer) {}
        InnerIterator(const SparseMatrix &m, Index outer)
            : m_row(const_cast<std::vector<Entry> *>(&m.m_data[static_cast<size_t>(outer)])), m_outer(outer)
        {
        }
        InnerIterator &operator++()
        {
            m_pos++;
            return *this;
        }
        operator bool() const { return m_pos < m_row->size(); }
        Index index() const { return (*m_row)[m_pos].first; }
        Index row() const { return m_outer; }
        Index col() const { return (*m_row)[m_pos].first; }
        S value() const { return (*m_row)[m_pos].second; }
        S &valueRef() { return (*m_row)[m_pos].second; }

    private:
        std::vector<Entry> *m_row;
        Index m_outer;
        size_t m_pos = 0;
    };

    template <int Mode> SparseMatrix triangularView() const
    {
        SparseMatrix out(m_rows, m_cols);
        for (Index r = 0; r < m_rows; r++)
            for (const Entry &e : m_data[static_cast<size_t>(r)])
            {
                bool keep = (Mode == StrictlyLower) ? (e.first < r)
                            : (Mode == Lower)       ? (e.first <= r)
                            : (Mode == StrictlyUpper) ? (e.first > r)
                                                      : (e.first >= r);
                if (keep) out.m_data[static_cast<size_t>(r)].push_back(e);
            }
        return out;
    }

    SparseMatrix transpose() const
    {
        SparseMatrix out(m_cols, m_rows);
        for (Index r = 0; r < m_rows; r++)
            for (const Entry &e : m_data[static_cast<size_t>(r)])
                out.m_data[static_cast<size_t>(e.first)].push_back(Entry(r, e.second));
        return out;
    }

    SparseMatrix operator-(const SparseMatrix &o) const
    {
        SparseMatrix out(*this);
        for (Index r = 0; r < o.m_rows; r++)
            for (const Entry &e : o.m_data[static_cast<size_t>(r)]) out.coeffRef(r, e.first) -= e.second;
        return out;
    }

    SparseMatrix operator*(const SparseMatrix &o) const
    {
        SparseMatrix out(m_rows, o.m_cols);
        for (Index r = 0; r < m_rows; r++)
            for (const Entry &a : m_data[static_cast<size_t>(r)])
                for (const Entry &b : o.m_data[static_cast<size_t>(a.first)])
                    out.coeffRef(r, b.first) += a.second * b.second;
        return out;
    }

    DenseVector<S> operator*(const DenseVector<S> &x) const
    {
        DenseVector<S> y(m_rows);
        for (Index r = 0; r < m_rows; r++)
        {
            S acc = S(0);
            for (const Entry &e : m_data[static_cast<size_t>(r)]) acc += e.second * x[e.first];
            y[r] = acc;
        }
        return y;
    }

    const std::vector<std::vector<Entry>> &rowLists() const { return m_data; }

private:
    template <int O2, typename SI2> void copyFrom(const SparseMatrix<S, O2, SI2> &o)
    {
        m_rows = o.rows();
        m_cols = o.cols();
        m_data = o.rowLists();
    }

    Index m_rows = 0;
    Index m_cols = 0;
    std::vector<std::vector<Entry>> m_data;
};

// Eigen 3.4.0 ConjugateGradient<MatrixType, UpLo, DiagonalPreconditioner>:
// restated from the published algorithm (Eigen/src/IterativeLinearSolvers/
// ConjugateGradient.h, conjugate_gradient()): self-adjoint view of the UpLo
// triangle, Jacobi preconditioner, x0 = 0, stop when |r|^2 < tol^2 |b|^2,
// maxIterations = 2 * cols.
template <typename MatrixType, int UpLo = Lower> class ConjugateGradient
{
public:
    typedef typename MatrixType::Scalar Scalar;
    typedef DenseVector<Scalar> Vector;

    ConjugateGradient() = default;

    ConjugateGradient &setTolerance(Scalar tol)
    {
        m_tolerance = tol;
        return *this;
    }
    ConjugateGradient &setMaxIterations(Index it)
    {
        m_maxIterations = it;
        return *this;
    }

    template <typename M2> ConjugateGradient &compute(const M2 &mat)
    {
        m_mat = mat;
        const Index n = m_mat.cols();
        m_invdiag.resize(n);
        for (Index j = 0; j < n; j++)
        {
            Scalar d = Scalar(0);
            bool found = false;
            for (const auto &e : m_mat.rowLists()[static_cast<size_t>(j)])
                if (e.first == j)
                {
                    d = e.second;
                    found = true;
                }
            m_invdiag[j] = (found && d != Scalar(0)) ? Scalar(1) / d : Scalar(1);
        }
        m_info = Success;
        m_isInitialized = true;
        return *this;
    }

    ComputationInfo info() const { return m_info; }
    Index iterations() const { return m_iterations; }
    Scalar error() const { return m_error; }

    Vector solve(const Vector &rhs) const
    {
        const Index n = m_mat.cols();
        Vector x(n);
        x.setZero();
        Index maxIters = m_maxIterations < 0 ? 2 * n : m_maxIterations;
        Scalar tol = m_tolerance;

        Vector residual(n);
        for (Index i = 0; i < n; i++) residual[i] = rhs[i];  // rhs - A*0
        Scalar rhsNorm2 = rhs.squaredNorm();
        if (rhsNorm2 == 0)
        {
            m_iterations = 0;
            m_error = 0;
            m_info = Success;
            return x;
        }
        const Scalar considerAsZero = (std::numeric_limits<Scalar>::min)();
        Scalar threshold = std::max(Scalar(tol * tol * rhsNorm2), considerAsZero);
        Scalar residualNorm2 = residual.squaredNorm();
        if (residualNorm2 < threshold)
        {
            m_iterations = 0;
            m_error = std::sqrt(residualNorm2 / rhsNorm2);
            m_info = Success;
            return x;
        }
        Vector p(n), z(n), tmp(n);
        for (Index i = 0; i < n; i++) p[i] = m_invdiag[i] * residual[i];
        Scalar absNew = residual.dot(p);
        Index i = 0;
        while (i < maxIters)
        {
            symProduct(p, tmp);
            Scalar alpha = absNew / p.dot(tmp);
            for (Index k = 0; k < n; k++) x[k] += alpha * p[k];
            for (Index k = 0; k < n; k++) residual[k] -= alpha * tmp[k];
            residualNorm2 = residual.squaredNorm();
            if (residualNorm2 < threshold) break;
            for (Index k = 0; k < n; k++) z[k] = m_invdiag[k] * residual[k];
            Scalar absOld = absNew;
            absNew = residual.dot(z);
            Scalar beta = absNew / absOld;
            for (Index k = 0; k < n; k++) p[k] = z[k] + beta * p[k];
            i++;
        }
        m_error = std::sqrt(residualNorm2 / rhsNorm2);
        m_iterations = i;
        m_info = m_error <= m_tolerance ? Success : NoConvergence;
        return x;
    }

private:
    // y = selfadjointView<UpLo>(A) * x
    void symProduct(const Vector &x, Vector &y) const
    {
        const Index n = m_mat.rows();
        y.setZero();
        for (Index r = 0; r < n; r++)
            for (const auto &e : m_mat.rowLists()[static_cast<size_t>(r)])
            {
                const Index c = e.first;
                const bool inTriangle = (UpLo == Upper) ? (c >= r) : (c <= r);
                if (!inTriangle) continue;
                y[r] += e.second * x[c];
                if (c != r) y[c] += e.second * x[r];
            }
    }

    SparseMatrix<Scalar, RowMajor> m_mat;
    Vector m_invdiag;
    Scalar m_tolerance = std::numeric_limits<Scalar>::epsilon();
    Index m_maxIterations = -1;
    mutable Index m_iterations = 0;
    mutable Scalar m_error = 0;
    mutable ComputationInfo m_info = Success;
    bool m_isInitialized = false;
};

}  // namespace Eigen

#endif  // FS2D_ORACLE_EIGEN_SHIM_H
