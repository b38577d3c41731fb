// LBFGSpp/SmallDense.h -- host-side dense algebra for the 2m x 2m "middle" matrices of L-BFGS-B.
//
// SmallMatrix: row-major dense storage shared by the host-side solvers.  SmallSolver: LU with partial pivoting for general
// square systems -- the stand-in for Eigen's PartialPivLU that the reference uses when it inverts the dense approximate
// Hessian (reference BFGSMat.h:205).  The symmetric indefinite middle matrices are factorised by BKLDLT.h.
#ifndef LBFGSPP_B200_SMALL_DENSE_H
#define LBFGSPP_B200_SMALL_DENSE_H

#include <cmath>
#include <stdexcept>
#include <utility>
#include <vector>

namespace LBFGSpp {

enum COMPUTATION_INFO { SUCCESSFUL = 0, NOT_COMPUTED, NUMERICAL_ISSUE };

template <typename Scalar>
class SmallMatrix
{
    int m_rows, m_cols;
    std::vector<Scalar> m_data;

public:
    SmallMatrix() : m_rows(0), m_cols(0) {}
    SmallMatrix(int rows, int cols, Scalar fill = Scalar(0)) : m_rows(rows), m_cols(cols), m_data(size_t(rows) * cols, fill) {}
    int rows() const { return m_rows; }
    int cols() const { return m_cols; }
    Scalar& operator()(int i, int j) { return m_data[size_t(i) * m_cols + j]; }
    const Scalar& operator()(int i, int j) const { return m_data[size_t(i) * m_cols + j]; }
    Scalar* data() { return m_data.data(); }
    const Scalar* data() const { return m_data.data(); }
    std::vector<Scalar> times(const std::vector<Scalar>& v) const
    {
        std::vector<Scalar> out(size_t(m_rows), Scalar(0));
        for (int i = 0; i < m_rows; i++)
        {
            Scalar acc = Scalar(0);
            for (int j = 0; j < m_cols; j++) acc += (*this)(i, j) * v[size_t(j)];
            out[size_t(i)] = acc;
        }
        return out;
    }
};

// General square solver; interface: compute / solve_inplace / solve / info.
template <typename Scalar>
class SmallSolver
{
    int m_n;
    SmallMatrix<Scalar> m_lu;
    std::vector<int> m_piv;
    int m_info;

public:
    SmallSolver() : m_n(0), m_info(NOT_COMPUTED) {}
    explicit SmallSolver(const SmallMatrix<Scalar>& a) : m_n(0), m_info(NOT_COMPUTED) { compute(a); }

    void compute(const SmallMatrix<Scalar>& a)
    {
        using std::abs;
        if (a.rows() != a.cols()) throw std::invalid_argument("SmallSolver: matrix must be square");
        m_n = a.rows();
        m_lu = a;
        m_piv.assign(size_t(m_n), 0);
        m_info = SUCCESSFUL;
        for (int k = 0; k < m_n; k++)
        {
            int best = k;
            for (int i = k + 1; i < m_n; i++)
                if (abs(m_lu(i, k)) > abs(m_lu(best, k))) best = i;
            m_piv[size_t(k)] = best;
            if (best != k)
                for (int j = 0; j < m_n; j++) std::swap(m_lu(k, j), m_lu(best, j));
            const Scalar pivot = m_lu(k, k);
            if (pivot == Scalar(0)) { m_info = NUMERICAL_ISSUE; continue; }
            for (int i = k + 1; i < m_n; i++)
            {
                const Scalar f = m_lu(i, k) / pivot;
                m_lu(i, k) = f;
                for (int j = k + 1; j < m_n; j++) m_lu(i, j) -= f * m_lu(k, j);
            }
        }
    }

    void solve_inplace(std::vector<Scalar>& b) const
    {
        if (m_info == NOT_COMPUTED) throw std::logic_error("SmallSolver: need to call compute() first");
        for (int k = 0; k < m_n; k++)
            if (m_piv[size_t(k)] != k) std::swap(b[size_t(k)], b[size_t(m_piv[size_t(k)])]);
        for (int i = 0; i < m_n; i++)
            for (int k = 0; k < i; k++) b[size_t(i)] -= m_lu(i, k) * b[size_t(k)];
        for (int i = m_n - 1; i >= 0; i--)
        {
            for (int k = i + 1; k < m_n; k++) b[size_t(i)] -= m_lu(i, k) * b[size_t(k)];
            b[size_t(i)] /= m_lu(i, i);
        }
    }
    std::vector<Scalar> solve(std::vector<Scalar> b) const
    {
        solve_inplace(b);
        return b;
    }
    int info() const { return m_info; }
};

}  // namespace LBFGSpp

#endif  // LBFGSPP_B200_SMALL_DENSE_H
