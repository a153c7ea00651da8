/*
 * bfmatrix.h - TEST INFRASTRUCTURE ONLY. Stand-in for the small part of FSL's MISCMATHS::BFMatrix that
 * inference_nlls.cc touches (a dense "full" matrix behind a virtual interface): Nrows / Ncols / Set / Peek and
 * SolveForx. FSL is not vendored in the reference tree; see armawrap/newmat.h in this directory.
 */
#ifndef FABBER_SHIM_BFMATRIX_H
#define FABBER_SHIM_BFMATRIX_H

#include "armawrap/newmat.h"
#include <boost/shared_ptr.hpp>

namespace MISCMATHS
{
enum MatrixType
{
    UNKNOWN,
    ASYM,
    SYM,
    SYM_POSDEF
};

class BFMatrix
{
public:
    virtual ~BFMatrix() {}
    virtual unsigned int Nrows() const = 0;
    virtual unsigned int Ncols() const = 0;
    virtual void Set(unsigned int r, unsigned int c, double v) = 0;
    virtual double Peek(unsigned int r, unsigned int c) const = 0;
    virtual NEWMAT::ReturnMatrix SolveForx(const NEWMAT::ColumnVector &b, MatrixType type, double tol, int miter) const = 0;
};

class FullBFMatrix : public BFMatrix
{
public:
    FullBFMatrix(unsigned int m, unsigned int n)
        : mp(new NEWMAT::Matrix(m, n))
    {
        *mp = 0.0;
    }
    unsigned int Nrows() const { return mp->Nrows(); }
    unsigned int Ncols() const { return mp->Ncols(); }
    void Set(unsigned int r, unsigned int c, double v) { (*mp)(r, c) = v; }
    double Peek(unsigned int r, unsigned int c) const { return (*mp)(r, c); }
    /* FSL: `ret = mp->i() * b` for a full matrix, whatever the type / tolerance arguments say */
    NEWMAT::ReturnMatrix SolveForx(const NEWMAT::ColumnVector &b, MatrixType, double, int) const
    {
        NEWMAT::ColumnVector ret;
        ret = mp->i() * b;
        ret.Release();
        return ret;
    }

private:
    boost::shared_ptr<NEWMAT::Matrix> mp;
};
} // namespace MISCMATHS
#endif
