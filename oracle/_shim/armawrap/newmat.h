/*
 * newmat.h - TEST INFRASTRUCTURE ONLY. A minimal, eager, dense stand-in for the NEWMAT API subset that
 * fabber_core's VB path uses (the real one is FSL's armawrap NEWMAT-over-Armadillo shim, which is not
 * vendored in the reference tree and not installed here). It exists so that the reference's OWN source
 * files can be compiled unchanged from /root/reference into oracle/_ref/libfabbercore_ref.so
 * (oracle/Makefile), which then pins the restated oracle and serves as the "reference" CPU baseline.
 *
 * Written from the NEWMAT documentation's public interface; 1-based element access; everything is a
 * dense row-major array, symmetric / diagonal types only constrain how values are stored on assignment.
 * Inverse and log-determinant are LU with partial pivoting (what LAPACK getrf does underneath Armadillo).
 * Nothing in the product includes this file.
 */
#ifndef FABBER_SHIM_NEWMAT_H
#define FABBER_SHIM_NEWMAT_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <exception>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

namespace NEWMAT
{
typedef double Real;

class Exception : public std::exception
{
public:
    explicit Exception(const std::string &m = "NEWMAT exception")
        : msg(m)
    {
    }
    virtual ~Exception() throw() {}
    virtual const char *what() const throw() { return msg.c_str(); }

private:
    std::string msg;
};
class SingularException : public Exception
{
public:
    SingularException()
        : Exception("matrix is singular")
    {
    }
};
class IncompatibleDimensionsException : public Exception
{
public:
    IncompatibleDimensionsException()
        : Exception("incompatible dimensions")
    {
    }
};

class LogAndSign
{
public:
    LogAndSign()
        : lv(0)
        , sg(1)
    {
    }
    LogAndSign(Real l, int s)
        : lv(l)
        , sg(s)
    {
    }
    Real LogValue() const { return lv; }
    int Sign() const { return sg; }
    Real Value() const { return sg * std::exp(lv); }

private:
    Real lv;
    int sg;
};

enum Shape
{
    GENERAL,
    SYMMETRIC,
    DIAGONAL,
    COLUMN,
    ROW
};

class Matrix;
class SubMatrixRef;
class GeneralMatrix;

/* NEWMAT list loading: `m << a << b << c;` stores consecutive elements in row order */
class ListLoader
{
public:
    ListLoader(GeneralMatrix &m_, size_t next_)
        : m(m_)
        , next(next_)
    {
    }
    ListLoader operator<<(Real v);

private:
    GeneralMatrix &m;
    size_t next;
};

/* dense base: every NEWMAT type used by fabber derives from this */
class GeneralMatrix
{
public:
    int nr, nc;
    std::vector<Real> a;
    Shape shape;

    GeneralMatrix(int r = 0, int c = 0, Shape s = GENERAL)
        : nr(r)
        , nc(c)
        , a((size_t)r * c, 0.0)
        , shape(s)
    {
    }
    virtual ~GeneralMatrix() {}
    int Nrows() const { return nr; }
    int Ncols() const { return nc; }
    int Storage() const { return nr * nc; }
    Real *Store() { return a.empty() ? 0 : &a[0]; }
    const Real *Store() const { return a.empty() ? 0 : &a[0]; }
    Real &at(int i, int j) { return a[(size_t)(i - 1) * nc + (j - 1)]; }
    Real at(int i, int j) const { return a[(size_t)(i - 1) * nc + (j - 1)]; }
    void check(int i, int j) const
    {
        if (i < 1 || j < 1 || i > nr || j > nc)
            throw Exception("index out of range");
    }
    /* element access: two indices for matrices, one for vectors / diagonal matrices */
    Real &operator()(int i, int j)
    {
        check(i, j);
        return at(i, j);
    }
    Real operator()(int i, int j) const
    {
        check(i, j);
        return at(i, j);
    }
    Real &operator()(int i)
    {
        if (shape == DIAGONAL)
        {
            check(i, i);
            return at(i, i);
        }
        if (nc == 1)
        {
            check(i, 1);
            return at(i, 1);
        }
        check(1, i);
        return at(1, i);
    }
    Real operator()(int i) const
    {
        if (shape == DIAGONAL)
        {
            check(i, i);
            return at(i, i);
        }
        if (nc == 1)
        {
            check(i, 1);
            return at(i, 1);
        }
        check(1, i);
        return at(1, i);
    }
    void set_size(int r, int c)
    {
        nr = r;
        nc = c;
        a.assign((size_t)r * c, 0.0);
    }
    /* a symmetric matrix keeps the lower triangle of whatever it is given; a diagonal one the diagonal */
    void conform()
    {
        if (shape == SYMMETRIC)
        {
            for (int i = 1; i <= nr; i++)
                for (int j = i + 1; j <= nc; j++)
                    at(i, j) = at(j, i);
        }
        else if (shape == DIAGONAL)
        {
            for (int i = 1; i <= nr; i++)
                for (int j = 1; j <= nc; j++)
                    if (i != j)
                        at(i, j) = 0.0;
        }
    }
    void fill(Real v)
    {
        if (shape == DIAGONAL)
        {
            std::fill(a.begin(), a.end(), 0.0);
            for (int i = 1; i <= nr; i++)
                at(i, i) = v;
        }
        else
            std::fill(a.begin(), a.end(), v);
    }
    void assign_from(const GeneralMatrix &m)
    {
        if (shape == COLUMN && m.nc != 1 && m.nr * m.nc != 0)
        {
            if (m.nr == 1)
            {
                nr = m.nc;
                nc = 1;
                a = m.a;
                return;
            }
            throw IncompatibleDimensionsException();
        }
        if (shape == ROW && m.nr != 1 && m.nr * m.nc != 0)
        {
            if (m.nc == 1)
            {
                nr = 1;
                nc = m.nr;
                a = m.a;
                return;
            }
            throw IncompatibleDimensionsException();
        }
        if ((shape == SYMMETRIC || shape == DIAGONAL) && m.nr != m.nc)
            throw IncompatibleDimensionsException();
        nr = m.nr;
        nc = m.nc;
        a = m.a;
        conform();
    }

    Matrix t() const;
    Matrix i() const;
    Real Trace() const
    {
        Real s = 0;
        for (int k = 1; k <= std::min(nr, nc); k++)
            s += at(k, k);
        return s;
    }
    Real AsScalar() const
    {
        if (nr != 1 || nc != 1)
            throw Exception("AsScalar: not 1x1");
        return a[0];
    }
    Real Sum() const
    {
        Real s = 0;
        for (size_t k = 0; k < a.size(); k++)
            s += a[k];
        return s;
    }
    Real SumSquare() const
    {
        Real s = 0;
        for (size_t k = 0; k < a.size(); k++)
            s += a[k] * a[k];
        return s;
    }
    Real SumAbsoluteValue() const
    {
        Real s = 0;
        for (size_t k = 0; k < a.size(); k++)
            s += std::fabs(a[k]);
        return s;
    }
    Real Maximum() const
    {
        if (a.empty())
            throw Exception("Maximum of empty matrix");
        Real m = -INFINITY;
        if (shape == DIAGONAL)
        {
            for (int k = 1; k <= nr; k++)
                m = std::max(m, at(k, k));
            return m;
        }
        for (size_t k = 0; k < a.size(); k++)
            m = std::max(m, a[k]);
        return m;
    }
    Real Minimum() const
    {
        if (a.empty())
            throw Exception("Minimum of empty matrix");
        Real m = INFINITY;
        if (shape == DIAGONAL)
        {
            for (int k = 1; k <= nr; k++)
                m = std::min(m, at(k, k));
            return m;
        }
        for (size_t k = 0; k < a.size(); k++)
            m = std::min(m, a[k]);
        return m;
    }
    Real MaximumAbsoluteValue() const
    {
        Real m = 0;
        for (size_t k = 0; k < a.size(); k++)
            m = std::max(m, std::fabs(a[k]));
        return m;
    }
    bool IsZero() const
    {
        for (size_t k = 0; k < a.size(); k++)
            if (a[k] != 0.0)
                return false;
        return true;
    }
    LogAndSign LogDeterminant() const;
    Real Determinant() const { return LogDeterminant().Value(); }
    Matrix AsRow() const;
    Matrix AsColumn() const;
    Matrix AsDiagonal() const;
    Matrix AsMatrix(int r, int c) const;

    SubMatrixRef SubMatrix(int r1, int r2, int c1, int c2);
    SubMatrixRef SymSubMatrix(int r1, int r2);
    SubMatrixRef Rows(int r1, int r2);
    SubMatrixRef Columns(int c1, int c2);
    SubMatrixRef Row(int r);
    SubMatrixRef Column(int c);
    Matrix SubMatrix(int r1, int r2, int c1, int c2) const;
    Matrix SymSubMatrix(int r1, int r2) const;
    Matrix Rows(int r1, int r2) const;
    Matrix Columns(int c1, int c2) const;
    Matrix Row(int r) const;
    Matrix Column(int c) const;
    void Release() {}
    void CleanUp() { set_size(0, 0); }
};

class Matrix : public GeneralMatrix
{
public:
    Matrix()
        : GeneralMatrix(0, 0, GENERAL)
    {
    }
    Matrix(int r, int c)
        : GeneralMatrix(r, c, GENERAL)
    {
    }
    Matrix(const GeneralMatrix &m)
        : GeneralMatrix(m.nr, m.nc, GENERAL)
    {
        a = m.a;
    }
    Matrix(const SubMatrixRef &s);
    Matrix &operator=(const GeneralMatrix &m)
    {
        nr = m.nr;
        nc = m.nc;
        a = m.a;
        return *this;
    }
    Matrix &operator=(const Matrix &m)
    {
        nr = m.nr;
        nc = m.nc;
        a = m.a;
        return *this;
    }
    Matrix &operator=(Real v)
    {
        fill(v);
        return *this;
    }
    Matrix &operator=(const SubMatrixRef &s);
    Matrix &operator<<(const SubMatrixRef &s);
    Matrix &operator<<(const GeneralMatrix &m) { return *this = m; }
    ListLoader operator<<(Real v) { return ListLoader(*this, 0) << v; }
    Matrix &operator<<(const Real *p)
    {
        for (size_t k = 0; k < a.size(); k++)
            a[k] = p[k];
        return *this;
    }
    void ReSize(int r, int c) { set_size(r, c); }
    void ReSize(const GeneralMatrix &m) { set_size(m.nr, m.nc); }
    Matrix &operator+=(const GeneralMatrix &m);
    Matrix &operator-=(const GeneralMatrix &m);
    Matrix &operator*=(Real v)
    {
        for (size_t k = 0; k < a.size(); k++)
            a[k] *= v;
        return *this;
    }
    Matrix &operator/=(Real v)
    {
        for (size_t k = 0; k < a.size(); k++)
            a[k] /= v;
        return *this;
    }
    Matrix &operator&=(const GeneralMatrix &m);
    Matrix &operator|=(const GeneralMatrix &m);
};
typedef Matrix ReturnMatrix;

class ColumnVector : public GeneralMatrix
{
public:
    ColumnVector()
        : GeneralMatrix(0, 1, COLUMN)
    {
    }
    explicit ColumnVector(int n)
        : GeneralMatrix(n, 1, COLUMN)
    {
    }
    ColumnVector(const GeneralMatrix &m)
        : GeneralMatrix(0, 1, COLUMN)
    {
        assign_from(m);
    }
    ColumnVector(const ColumnVector &m)
        : GeneralMatrix(m.nr, 1, COLUMN)
    {
        a = m.a;
    }
    ColumnVector(const SubMatrixRef &s);
    ColumnVector &operator=(const GeneralMatrix &m)
    {
        assign_from(m);
        return *this;
    }
    ColumnVector &operator=(const ColumnVector &m)
    {
        nr = m.nr;
        nc = 1;
        a = m.a;
        return *this;
    }
    ColumnVector &operator=(Real v)
    {
        fill(v);
        return *this;
    }
    ColumnVector &operator=(const SubMatrixRef &s);
    ColumnVector &operator<<(const SubMatrixRef &s);
    ColumnVector &operator<<(const GeneralMatrix &m)
    {
        assign_from(m);
        return *this;
    }
    ListLoader operator<<(Real v) { return ListLoader(*this, 0) << v; }
    ColumnVector &operator<<(const Real *p)
    {
        for (size_t k = 0; k < a.size(); k++)
            a[k] = p[k];
        return *this;
    }
    void ReSize(int n) { set_size(n, 1); }
    ColumnVector &operator+=(const GeneralMatrix &m);
    ColumnVector &operator-=(const GeneralMatrix &m);
    ColumnVector &operator*=(Real v)
    {
        for (size_t k = 0; k < a.size(); k++)
            a[k] *= v;
        return *this;
    }
    ColumnVector &operator/=(Real v)
    {
        for (size_t k = 0; k < a.size(); k++)
            a[k] /= v;
        return *this;
    }
    ColumnVector &operator&=(const GeneralMatrix &m);
};

class RowVector : public GeneralMatrix
{
public:
    RowVector()
        : GeneralMatrix(1, 0, ROW)
    {
    }
    explicit RowVector(int n)
        : GeneralMatrix(1, n, ROW)
    {
    }
    RowVector(const GeneralMatrix &m)
        : GeneralMatrix(1, 0, ROW)
    {
        assign_from(m);
    }
    RowVector(const RowVector &m)
        : GeneralMatrix(1, m.nc, ROW)
    {
        a = m.a;
    }
    RowVector(const SubMatrixRef &s);
    RowVector &operator=(const GeneralMatrix &m)
    {
        assign_from(m);
        return *this;
    }
    RowVector &operator=(const RowVector &m)
    {
        nr = 1;
        nc = m.nc;
        a = m.a;
        return *this;
    }
    RowVector &operator=(Real v)
    {
        fill(v);
        return *this;
    }
    RowVector &operator=(const SubMatrixRef &s);
    RowVector &operator<<(const SubMatrixRef &s);
    RowVector &operator<<(const GeneralMatrix &m)
    {
        assign_from(m);
        return *this;
    }
    ListLoader operator<<(Real v) { return ListLoader(*this, 0) << v; }
    void ReSize(int n) { set_size(1, n); }
};

class SymmetricMatrix : public GeneralMatrix
{
public:
    SymmetricMatrix()
        : GeneralMatrix(0, 0, SYMMETRIC)
    {
    }
    explicit SymmetricMatrix(int n)
        : GeneralMatrix(n, n, SYMMETRIC)
    {
    }
    SymmetricMatrix(const GeneralMatrix &m)
        : GeneralMatrix(0, 0, SYMMETRIC)
    {
        assign_from(m);
    }
    SymmetricMatrix(const SymmetricMatrix &m)
        : GeneralMatrix(m.nr, m.nc, SYMMETRIC)
    {
        a = m.a;
    }
    SymmetricMatrix(const SubMatrixRef &s);
    SymmetricMatrix &operator=(const GeneralMatrix &m)
    {
        assign_from(m);
        return *this;
    }
    SymmetricMatrix &operator=(const SymmetricMatrix &m)
    {
        nr = m.nr;
        nc = m.nc;
        a = m.a;
        return *this;
    }
    SymmetricMatrix &operator=(Real v)
    {
        fill(v);
        return *this;
    }
    SymmetricMatrix &operator=(const SubMatrixRef &s);
    SymmetricMatrix &operator<<(const SubMatrixRef &s);
    SymmetricMatrix &operator<<(const GeneralMatrix &m)
    {
        assign_from(m);
        return *this;
    }
    ListLoader operator<<(Real v) { return ListLoader(*this, 0) << v; }
    void ReSize(int n) { set_size(n, n); }
    /* symmetric element access writes both triangles */
    class ElemRef
    {
    public:
        ElemRef(SymmetricMatrix &m_, int i_, int j_)
            : m(m_)
            , i(i_)
            , j(j_)
        {
        }
        operator Real() const { return m.at(i, j); }
        ElemRef &operator=(Real v)
        {
            m.at(i, j) = v;
            m.at(j, i) = v;
            return *this;
        }
        ElemRef &operator=(const ElemRef &o) { return *this = (Real)o; }
        ElemRef &operator+=(Real v) { return *this = m.at(i, j) + v; }
        ElemRef &operator-=(Real v) { return *this = m.at(i, j) - v; }
        ElemRef &operator*=(Real v) { return *this = m.at(i, j) * v; }
        ElemRef &operator/=(Real v) { return *this = m.at(i, j) / v; }

    private:
        SymmetricMatrix &m;
        int i, j;
    };
    ElemRef operator()(int i, int j)
    {
        check(i, j);
        return ElemRef(*this, i, j);
    }
    Real operator()(int i, int j) const
    {
        check(i, j);
        return at(i, j);
    }
    SymmetricMatrix &operator+=(const GeneralMatrix &m);
    SymmetricMatrix &operator-=(const GeneralMatrix &m);
    SymmetricMatrix &operator*=(Real v)
    {
        for (size_t k = 0; k < a.size(); k++)
            a[k] *= v;
        return *this;
    }
};

class DiagonalMatrix : public GeneralMatrix
{
public:
    DiagonalMatrix()
        : GeneralMatrix(0, 0, DIAGONAL)
    {
    }
    explicit DiagonalMatrix(int n)
        : GeneralMatrix(n, n, DIAGONAL)
    {
    }
    DiagonalMatrix(const GeneralMatrix &m)
        : GeneralMatrix(0, 0, DIAGONAL)
    {
        assign_from(m);
    }
    DiagonalMatrix(const DiagonalMatrix &m)
        : GeneralMatrix(m.nr, m.nc, DIAGONAL)
    {
        a = m.a;
    }
    DiagonalMatrix &operator=(const GeneralMatrix &m)
    {
        assign_from(m);
        return *this;
    }
    DiagonalMatrix &operator=(const DiagonalMatrix &m)
    {
        nr = m.nr;
        nc = m.nc;
        a = m.a;
        return *this;
    }
    DiagonalMatrix &operator=(Real v)
    {
        fill(v);
        return *this;
    }
    DiagonalMatrix &operator=(const SubMatrixRef &s);
    DiagonalMatrix &operator<<(const SubMatrixRef &s);
    DiagonalMatrix &operator<<(const GeneralMatrix &m)
    {
        assign_from(m);
        return *this;
    }
    ListLoader operator<<(Real v) { return ListLoader(*this, 0) << v; }
    void ReSize(int n) { set_size(n, n); }
    DiagonalMatrix &operator+=(const GeneralMatrix &m);
    DiagonalMatrix &operator-=(const GeneralMatrix &m);
};

class IdentityMatrix : public GeneralMatrix
{
public:
    explicit IdentityMatrix(int n = 0)
        : GeneralMatrix(n, n, DIAGONAL)
    {
        for (int i = 1; i <= n; i++)
            at(i, i) = 1.0;
    }
    void ReSize(int n)
    {
        set_size(n, n);
        for (int i = 1; i <= n; i++)
            at(i, i) = 1.0;
    }
};

/* assignable window on a matrix: m.Row(i) = ..., m.Column(j) = ..., m.SubMatrix(..) << ... */
class SubMatrixRef
{
public:
    GeneralMatrix &m;
    int r1, r2, c1, c2;
    SubMatrixRef(GeneralMatrix &m_, int r1_, int r2_, int c1_, int c2_)
        : m(m_)
        , r1(r1_)
        , r2(r2_)
        , c1(c1_)
        , c2(c2_)
    {
        if (r1 < 1 || c1 < 1 || r2 > m.nr || c2 > m.nc || r2 < r1 - 1 || c2 < c1 - 1)
            throw Exception("submatrix out of range");
    }
    int Nrows() const { return r2 - r1 + 1; }
    int Ncols() const { return c2 - c1 + 1; }
    Matrix value() const;
    void assign(const GeneralMatrix &v)
    {
        int nr = Nrows(), nc = Ncols();
        if (v.nr * v.nc != nr * nc)
            throw IncompatibleDimensionsException();
        /* same number of elements: accept a row for a column and vice versa, like NEWMAT's << */
        size_t k = 0;
        for (int i = 0; i < nr; i++)
            for (int j = 0; j < nc; j++)
                m.at(r1 + i, c1 + j) = v.a[k++];
        m.conform();
    }
    SubMatrixRef &operator=(const GeneralMatrix &v)
    {
        assign(v);
        return *this;
    }
    SubMatrixRef &operator=(const SubMatrixRef &v);
    SubMatrixRef &operator<<(const GeneralMatrix &v)
    {
        assign(v);
        return *this;
    }
    SubMatrixRef &operator=(Real x)
    {
        for (int i = r1; i <= r2; i++)
            for (int j = c1; j <= c2; j++)
                m.at(i, j) = x;
        m.conform();
        return *this;
    }
    SubMatrixRef &operator<<(Real x) { return *this = x; }
    SubMatrixRef &operator+=(const GeneralMatrix &v);
    SubMatrixRef &operator-=(const GeneralMatrix &v);
    SubMatrixRef &operator*=(Real x)
    {
        for (int i = r1; i <= r2; i++)
            for (int j = c1; j <= c2; j++)
                m.at(i, j) *= x;
        return *this;
    }
    SubMatrixRef &operator/=(Real x)
    {
        for (int i = r1; i <= r2; i++)
            for (int j = c1; j <= c2; j++)
                m.at(i, j) /= x;
        return *this;
    }
    /* read-only conveniences so a window can be used like a matrix in expressions */
    Matrix t() const;
    Matrix i() const;
    Real Sum() const;
    Real SumSquare() const;
    Real Maximum() const;
    Real Minimum() const;
    Real MaximumAbsoluteValue() const;
    Real AsScalar() const;
    Real Trace() const;
    Matrix AsRow() const;
    Matrix AsColumn() const;
    Matrix AsDiagonal() const;
    Real operator()(int i) const;
    Real operator()(int i, int j) const { return m.at(r1 + i - 1, c1 + j - 1); }
    /* nested windows are read-only values */
    Matrix Rows(int a, int b) const;
    Matrix Columns(int a, int b) const;
    Matrix Row(int a) const;
    Matrix Column(int a) const;
    Matrix SubMatrix(int a, int b, int c, int d) const;
};

inline ListLoader ListLoader::operator<<(Real v)
{
    if (next >= m.a.size())
        throw Exception("list loading: too many values");
    if (m.shape == DIAGONAL)
        m.at((int)next + 1, (int)next + 1) = v;
    else
        m.a[next] = v;
    return ListLoader(m, next + 1);
}

inline Matrix SubMatrixRef::value() const
{
    Matrix out(Nrows(), Ncols());
    for (int i = 0; i < out.nr; i++)
        for (int j = 0; j < out.nc; j++)
            out.at(i + 1, j + 1) = m.at(r1 + i, c1 + j);
    return out;
}
inline SubMatrixRef &SubMatrixRef::operator=(const SubMatrixRef &v)
{
    assign(v.value());
    return *this;
}
inline Matrix::Matrix(const SubMatrixRef &s)
    : GeneralMatrix(0, 0, GENERAL)
{
    *this = s.value();
}
inline ColumnVector::ColumnVector(const SubMatrixRef &s)
    : GeneralMatrix(0, 1, COLUMN)
{
    assign_from(s.value());
}
inline RowVector::RowVector(const SubMatrixRef &s)
    : GeneralMatrix(1, 0, ROW)
{
    assign_from(s.value());
}
inline SymmetricMatrix::SymmetricMatrix(const SubMatrixRef &s)
    : GeneralMatrix(0, 0, SYMMETRIC)
{
    assign_from(s.value());
}

inline SubMatrixRef GeneralMatrix::SubMatrix(int r1, int r2, int c1, int c2) { return SubMatrixRef(*this, r1, r2, c1, c2); }
inline SubMatrixRef GeneralMatrix::SymSubMatrix(int r1, int r2) { return SubMatrixRef(*this, r1, r2, r1, r2); }
inline SubMatrixRef GeneralMatrix::Rows(int r1, int r2) { return SubMatrixRef(*this, r1, r2, 1, nc); }
inline SubMatrixRef GeneralMatrix::Columns(int c1, int c2) { return SubMatrixRef(*this, 1, nr, c1, c2); }
inline SubMatrixRef GeneralMatrix::Row(int r) { return SubMatrixRef(*this, r, r, 1, nc); }
inline SubMatrixRef GeneralMatrix::Column(int c) { return SubMatrixRef(*this, 1, nr, c, c); }
inline Matrix GeneralMatrix::SubMatrix(int r1, int r2, int c1, int c2) const
{
    return SubMatrixRef(const_cast<GeneralMatrix &>(*this), r1, r2, c1, c2).value();
}
inline Matrix GeneralMatrix::SymSubMatrix(int r1, int r2) const { return SubMatrix(r1, r2, r1, r2); }
inline Matrix GeneralMatrix::Rows(int r1, int r2) const { return SubMatrix(r1, r2, 1, nc); }
inline Matrix GeneralMatrix::Columns(int c1, int c2) const { return SubMatrix(1, nr, c1, c2); }
inline Matrix GeneralMatrix::Row(int r) const { return SubMatrix(r, r, 1, nc); }
inline Matrix GeneralMatrix::Column(int c) const { return SubMatrix(1, nr, c, c); }

inline Matrix GeneralMatrix::t() const
{
    Matrix out(nc, nr);
    for (int i = 1; i <= nr; i++)
        for (int j = 1; j <= nc; j++)
            out.at(j, i) = at(i, j);
    return out;
}
inline Matrix GeneralMatrix::AsRow() const
{
    Matrix out(1, nr * nc);
    out.a = a;
    return out;
}
inline Matrix GeneralMatrix::AsColumn() const
{
    Matrix out(nr * nc, 1);
    if (shape == SYMMETRIC)
    {
        /* NEWMAT: AsColumn of a SymmetricMatrix lists the stored lower triangle by rows */
        Matrix tri(nr * (nr + 1) / 2, 1);
        size_t k = 0;
        for (int i = 1; i <= nr; i++)
            for (int j = 1; j <= i; j++)
                tri.a[k++] = at(i, j);
        return tri;
    }
    out.a = a;
    return out;
}
inline Matrix GeneralMatrix::AsDiagonal() const
{
    int n = nr * nc;
    Matrix out(n, n);
    for (int i = 0; i < n; i++)
        out.at(i + 1, i + 1) = a[i];
    return out;
}
inline Matrix GeneralMatrix::AsMatrix(int r, int c) const
{
    if (r * c != nr * nc)
        throw IncompatibleDimensionsException();
    Matrix out(r, c);
    out.a = a;
    return out;
}

/* ---- LU with partial pivoting ------------------------------------------------------------------- */
struct LUFactor
{
    int n;
    std::vector<Real> lu;
    std::vector<int> piv;
    int sign;
    bool singular;
    explicit LUFactor(const GeneralMatrix &A)
        : n(A.nr)
        , lu(A.a)
        , piv(A.nr)
        , sign(1)
        , singular(false)
    {
        if (A.nr != A.nc)
            throw Exception("LU of a non-square matrix");
        for (int k = 0; k < n; k++)
        {
            int p = k;
            Real best = std::fabs(lu[(size_t)k * n + k]);
            for (int i = k + 1; i < n; i++)
                if (std::fabs(lu[(size_t)i * n + k]) > best)
                {
                    best = std::fabs(lu[(size_t)i * n + k]);
                    p = i;
                }
            piv[k] = p;
            if (!(best > 0.0) || !std::isfinite(best))
            {
                singular = true;
                continue;
            }
            if (p != k)
            {
                for (int j = 0; j < n; j++)
                    std::swap(lu[(size_t)k * n + j], lu[(size_t)p * n + j]);
                sign = -sign;
            }
            for (int i = k + 1; i < n; i++)
            {
                lu[(size_t)i * n + k] /= lu[(size_t)k * n + k];
                const Real f = lu[(size_t)i * n + k];
                for (int j = k + 1; j < n; j++)
                    lu[(size_t)i * n + j] -= f * lu[(size_t)k * n + j];
            }
        }
    }
};

inline Matrix GeneralMatrix::i() const
{
    for (size_t k = 0; k < a.size(); k++)
        if (!std::isfinite(a[k]))
            throw SingularException();
    LUFactor f(*this);
    if (f.singular)
        throw SingularException();
    const int n = nr;
    Matrix inv(n, n);
    std::vector<Real> b(n);
    for (int col = 0; col < n; col++)
    {
        std::fill(b.begin(), b.end(), 0.0);
        b[col] = 1.0;
        for (int k = 0; k < n; k++)
            if (f.piv[k] != k)
                std::swap(b[k], b[f.piv[k]]);
        for (int r = 1; r < n; r++)
        {
            Real s = b[r];
            for (int j = 0; j < r; j++)
                s -= f.lu[(size_t)r * n + j] * b[j];
            b[r] = s;
        }
        for (int r = n - 1; r >= 0; r--)
        {
            Real s = b[r];
            for (int j = r + 1; j < n; j++)
                s -= f.lu[(size_t)r * n + j] * b[j];
            b[r] = s / f.lu[(size_t)r * n + r];
        }
        for (int r = 0; r < n; r++)
            inv.at(r + 1, col + 1) = b[r];
    }
    return inv;
}

inline LogAndSign GeneralMatrix::LogDeterminant() const
{
    LUFactor f(*this);
    Real lv = 0;
    int sg = f.sign;
    for (int k = 0; k < nr; k++)
    {
        const Real d = f.lu[(size_t)k * nr + k];
        if (d == 0.0 || f.singular)
            return LogAndSign(-INFINITY, 0);
        if (d < 0)
            sg = -sg;
        lv += std::log(std::fabs(d));
    }
    return LogAndSign(lv, sg);
}

/* ---- arithmetic (eager; results are general matrices) ------------------------------------------------ */
inline void same_size(const GeneralMatrix &A, const GeneralMatrix &B)
{
    if (A.nr != B.nr || A.nc != B.nc)
        throw IncompatibleDimensionsException();
}
inline Matrix operator+(const GeneralMatrix &A, const GeneralMatrix &B)
{
    same_size(A, B);
    Matrix C(A.nr, A.nc);
    for (size_t k = 0; k < C.a.size(); k++)
        C.a[k] = A.a[k] + B.a[k];
    return C;
}
inline Matrix operator-(const GeneralMatrix &A, const GeneralMatrix &B)
{
    same_size(A, B);
    Matrix C(A.nr, A.nc);
    for (size_t k = 0; k < C.a.size(); k++)
        C.a[k] = A.a[k] - B.a[k];
    return C;
}
inline Matrix operator-(const GeneralMatrix &A)
{
    Matrix C(A.nr, A.nc);
    for (size_t k = 0; k < C.a.size(); k++)
        C.a[k] = -A.a[k];
    return C;
}
inline Matrix operator*(const GeneralMatrix &A, const GeneralMatrix &B)
{
    if (A.nc != B.nr)
        throw IncompatibleDimensionsException();
    Matrix C(A.nr, B.nc);
    for (int i = 0; i < A.nr; i++)
        for (int j = 0; j < B.nc; j++)
        {
            Real s = 0;
            for (int k = 0; k < A.nc; k++)
                s += A.a[(size_t)i * A.nc + k] * B.a[(size_t)k * B.nc + j];
            C.a[(size_t)i * B.nc + j] = s;
        }
    return C;
}
inline Matrix operator*(const GeneralMatrix &A, Real v)
{
    Matrix C(A.nr, A.nc);
    for (size_t k = 0; k < C.a.size(); k++)
        C.a[k] = A.a[k] * v;
    return C;
}
inline Matrix operator*(Real v, const GeneralMatrix &A) { return A * v; }
inline Matrix operator/(const GeneralMatrix &A, Real v)
{
    Matrix C(A.nr, A.nc);
    for (size_t k = 0; k < C.a.size(); k++)
        C.a[k] = A.a[k] / v;
    return C;
}
inline Matrix operator+(const GeneralMatrix &A, Real v)
{
    Matrix C(A.nr, A.nc);
    for (size_t k = 0; k < C.a.size(); k++)
        C.a[k] = A.a[k] + v;
    return C;
}
inline Matrix operator-(const GeneralMatrix &A, Real v) { return A + (-v); }
inline Matrix SP(const GeneralMatrix &A, const GeneralMatrix &B)
{
    same_size(A, B);
    Matrix C(A.nr, A.nc);
    for (size_t k = 0; k < C.a.size(); k++)
        C.a[k] = A.a[k] * B.a[k];
    return C;
}
/* vertical / horizontal concatenation */
inline Matrix operator&(const GeneralMatrix &A, const GeneralMatrix &B)
{
    if (A.nr * A.nc == 0)
        return Matrix(B);
    if (B.nr * B.nc == 0)
        return Matrix(A);
    if (A.nc != B.nc)
        throw IncompatibleDimensionsException();
    Matrix C(A.nr + B.nr, A.nc);
    std::copy(A.a.begin(), A.a.end(), C.a.begin());
    std::copy(B.a.begin(), B.a.end(), C.a.begin() + A.a.size());
    return C;
}
inline Matrix operator|(const GeneralMatrix &A, const GeneralMatrix &B)
{
    if (A.nr * A.nc == 0)
        return Matrix(B);
    if (B.nr * B.nc == 0)
        return Matrix(A);
    if (A.nr != B.nr)
        throw IncompatibleDimensionsException();
    Matrix C(A.nr, A.nc + B.nc);
    for (int i = 1; i <= A.nr; i++)
    {
        for (int j = 1; j <= A.nc; j++)
            C.at(i, j) = A.at(i, j);
        for (int j = 1; j <= B.nc; j++)
            C.at(i, A.nc + j) = B.at(i, j);
    }
    return C;
}
inline bool operator==(const GeneralMatrix &A, const GeneralMatrix &B)
{
    if (A.nr != B.nr || A.nc != B.nc)
        return false;
    for (size_t k = 0; k < A.a.size(); k++)
        if (!(A.a[k] == B.a[k]))
            return false;
    return true;
}
inline bool operator!=(const GeneralMatrix &A, const GeneralMatrix &B) { return !(A == B); }

/* windows in expressions */
#define FAB_SHIM_BIN(op)                                                                                      \
    inline Matrix operator op(const SubMatrixRef &A, const GeneralMatrix &B) { return A.value() op B; }        \
    inline Matrix operator op(const GeneralMatrix &A, const SubMatrixRef &B) { return A op B.value(); }        \
    inline Matrix operator op(const SubMatrixRef &A, const SubMatrixRef &B) { return A.value() op B.value(); }
FAB_SHIM_BIN(+)
FAB_SHIM_BIN(-)
FAB_SHIM_BIN(*)
FAB_SHIM_BIN(&)
FAB_SHIM_BIN(|)
#undef FAB_SHIM_BIN
inline Matrix operator*(const SubMatrixRef &A, Real v) { return A.value() * v; }
inline Matrix operator*(Real v, const SubMatrixRef &A) { return A.value() * v; }
inline Matrix operator/(const SubMatrixRef &A, Real v) { return A.value() / v; }
inline Matrix operator-(const SubMatrixRef &A) { return -A.value(); }
inline bool operator==(const SubMatrixRef &A, const GeneralMatrix &B) { return A.value() == B; }
inline Matrix SP(const SubMatrixRef &A, const SubMatrixRef &B) { return SP(A.value(), B.value()); }
inline Matrix SP(const SubMatrixRef &A, const GeneralMatrix &B) { return SP(A.value(), B); }
inline Matrix SP(const GeneralMatrix &A, const SubMatrixRef &B) { return SP(A, B.value()); }

inline Matrix SubMatrixRef::t() const { return value().t(); }
inline Matrix SubMatrixRef::i() const { return value().i(); }
inline Real SubMatrixRef::Sum() const { return value().Sum(); }
inline Real SubMatrixRef::SumSquare() const { return value().SumSquare(); }
inline Real SubMatrixRef::Maximum() const { return value().Maximum(); }
inline Real SubMatrixRef::Minimum() const { return value().Minimum(); }
inline Real SubMatrixRef::MaximumAbsoluteValue() const { return value().MaximumAbsoluteValue(); }
inline Real SubMatrixRef::AsScalar() const { return value().AsScalar(); }
inline Real SubMatrixRef::Trace() const { return value().Trace(); }
inline Matrix SubMatrixRef::AsRow() const { return value().AsRow(); }
inline Matrix SubMatrixRef::AsColumn() const { return value().AsColumn(); }
inline Matrix SubMatrixRef::AsDiagonal() const { return value().AsDiagonal(); }
inline Real SubMatrixRef::operator()(int i) const
{
    Matrix v = value();
    return static_cast<const GeneralMatrix &>(v)(i);
}
inline Matrix SubMatrixRef::Rows(int a, int b) const { return static_cast<const GeneralMatrix &>(value()).Rows(a, b); }
inline Matrix SubMatrixRef::Columns(int a, int b) const { return static_cast<const GeneralMatrix &>(value()).Columns(a, b); }
inline Matrix SubMatrixRef::Row(int a) const { return static_cast<const GeneralMatrix &>(value()).Row(a); }
inline Matrix SubMatrixRef::Column(int a) const { return static_cast<const GeneralMatrix &>(value()).Column(a); }
inline Matrix SubMatrixRef::SubMatrix(int a, int b, int c, int d) const
{
    return static_cast<const GeneralMatrix &>(value()).SubMatrix(a, b, c, d);
}
#define FAB_SHIM_FROM_REF(T)                                                  \
    inline T &T::operator=(const SubMatrixRef &s) { return *this = s.value(); } \
    inline T &T::operator<<(const SubMatrixRef &s) { return *this = s.value(); }
FAB_SHIM_FROM_REF(Matrix)
FAB_SHIM_FROM_REF(ColumnVector)
FAB_SHIM_FROM_REF(RowVector)
FAB_SHIM_FROM_REF(SymmetricMatrix)
FAB_SHIM_FROM_REF(DiagonalMatrix)
#undef FAB_SHIM_FROM_REF
inline SubMatrixRef &SubMatrixRef::operator+=(const GeneralMatrix &v)
{
    assign(value() + v);
    return *this;
}
inline SubMatrixRef &SubMatrixRef::operator-=(const GeneralMatrix &v)
{
    assign(value() - v);
    return *this;
}

#define FAB_SHIM_COMPOUND(T)                                      \
    inline T &T::operator+=(const GeneralMatrix &m)               \
    {                                                             \
        same_size(*this, m);                                      \
        for (size_t k = 0; k < a.size(); k++)                     \
            a[k] += m.a[k];                                       \
        conform();                                                \
        return *this;                                             \
    }                                                             \
    inline T &T::operator-=(const GeneralMatrix &m)               \
    {                                                             \
        same_size(*this, m);                                      \
        for (size_t k = 0; k < a.size(); k++)                     \
            a[k] -= m.a[k];                                       \
        conform();                                                \
        return *this;                                             \
    }
FAB_SHIM_COMPOUND(Matrix)
FAB_SHIM_COMPOUND(ColumnVector)
FAB_SHIM_COMPOUND(SymmetricMatrix)
FAB_SHIM_COMPOUND(DiagonalMatrix)
#undef FAB_SHIM_COMPOUND
inline Matrix &Matrix::operator&=(const GeneralMatrix &m) { return *this = (*this & m); }
inline Matrix &Matrix::operator|=(const GeneralMatrix &m) { return *this = (*this | m); }
inline ColumnVector &ColumnVector::operator&=(const GeneralMatrix &m)
{
    Matrix tmp = static_cast<const GeneralMatrix &>(*this) & m;
    assign_from(tmp);
    return *this;
}

inline std::ostream &operator<<(std::ostream &os, const GeneralMatrix &m)
{
    for (int i = 1; i <= m.nr; i++)
    {
        for (int j = 1; j <= m.nc; j++)
            os << m.at(i, j) << " ";
        os << "\n";
    }
    return os;
}
inline std::ostream &operator<<(std::ostream &os, const SubMatrixRef &m) { return os << m.value(); }

} // namespace NEWMAT

#endif
