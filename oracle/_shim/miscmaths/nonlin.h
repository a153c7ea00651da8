/*
 * nonlin.h - TEST INFRASTRUCTURE ONLY. Stand-in for FSL's MISCMATHS::nonlin (miscmaths/nonlin.{h,cpp}), the
 * optimiser behind the reference's --method=nlls. FSL is not vendored in the reference tree and not installed here;
 * this is the builder's restatement of the published Levenberg(-Marquardt) driver `levmar` - the SAME restatement as
 * oracle/vb_oracle.cc::nlls_levmar, which is what the reference's golden test/outdata_linear_nlls pins
 * (tests/test_oracle_golden.py). With it the reference's OWN inference_nlls.cc (cost function, gradient, Hessian,
 * masked rows, precision = J'J / mse, 1e-6 floor, exception path) can run here and be compared with the oracle's
 * restatement of those parts (tests/test_reference_build.py); it is NOT an independent check of the optimiser.
 */
#ifndef FABBER_SHIM_NONLIN_H
#define FABBER_SHIM_NONLIN_H

#include <cmath>
#include <vector>

#include "armawrap/newmat.h"
#include "miscmaths/bfmatrix.h"
#include <boost/shared_ptr.hpp>

namespace MISCMATHS
{
enum NLMethod
{
    NL_VM,
    NL_CG,
    NL_SCG,
    NL_LM
};
enum LMType
{
    LM_L,
    LM_LM
};
enum NonlinOut
{
    NL_UNDEFINED,
    NL_MAXITER,
    NL_LM_MAXITER,
    NL_PARCONV,
    NL_GRADCONV,
    NL_CFCONV,
    NL_LCONV
};

class NonlinCF
{
public:
    virtual ~NonlinCF() {}
    virtual double cf(const NEWMAT::ColumnVector &p) const = 0;
    virtual NEWMAT::ReturnMatrix grad(const NEWMAT::ColumnVector &p) const = 0;
    virtual boost::shared_ptr<BFMatrix> hess(const NEWMAT::ColumnVector &p, boost::shared_ptr<BFMatrix> iptr) const = 0;
};

/* the state and the defaults nonlin reads: 200 accepted steps, lambda 0.1, lambda limit 1e20, fractional cost
 * tolerance 1e-8; NL_LM starts as Levenberg-Marquardt until SetGaussNewtonType(LM_L) */
class NonlinParam
{
public:
    NonlinParam(int npar, NLMethod mtd)
        : m_npar(npar)
        , m_mtd(mtd)
        , m_gntype(LM_LM)
        , m_maxiter(200)
        , m_niter(0)
        , m_lambda(0.1)
        , m_ltol(1.0e20)
        , m_cftol(1.0e-8)
        , m_cf(0.0)
        , m_logpar(false)
        , m_logcf(false)
        , m_status(NL_UNDEFINED)
    {
        m_par.ReSize(npar);
        m_par = 0.0;
    }
    int NPar() const { return m_npar; }
    NLMethod Method() const { return m_mtd; }
    LMType GaussNewtonType() const { return m_gntype; }
    void SetGaussNewtonType(LMType t) { m_gntype = t; }
    void SetStartingEstimate(const NEWMAT::ColumnVector &p) { m_par = p; }
    void LogPar(bool f) { m_logpar = f; }
    void LogCF(bool f) { m_logcf = f; }
    const NEWMAT::ColumnVector &Par() const { return m_par; }
    double CF() const { return m_cf; }
    double Lambda() const { return m_lambda; }
    double LambdaConvergenceCriterion() const { return m_ltol; }
    double FractionalCFTolerance() const { return m_cftol; }
    double EquationSolverTol() const { return 1.0e-3; }
    int EquationSolverMaxIter() const { return 200; }
    int NIter() const { return m_niter; }
    NonlinOut Status() const { return m_status; }
    const std::vector<double> &CFHistory() const { return m_cfhist; }
    const std::vector<NEWMAT::ColumnVector> &ParHistory() const { return m_parhist; }
    /* state is mutable: nonlin takes the parameters by const reference, as FSL's does */
    void SetPar(const NEWMAT::ColumnVector &p) const
    {
        m_par = p;
        if (m_logpar)
            m_parhist.push_back(p);
    }
    void SetCF(double cf) const
    {
        m_cf = cf;
        if (m_logcf)
            m_cfhist.push_back(cf);
    }
    void SetLambda(double l) const { m_lambda = l; }
    void SetStatus(NonlinOut s) const { m_status = s; }
    bool NextIter(bool success = true) const
    {
        if (success && m_niter++ >= m_maxiter)
            return false;
        return true;
    }

private:
    int m_npar;
    NLMethod m_mtd;
    LMType m_gntype;
    int m_maxiter;
    mutable int m_niter;
    mutable double m_lambda;
    double m_ltol, m_cftol;
    mutable double m_cf;
    bool m_logpar, m_logcf;
    mutable NonlinOut m_status;
    mutable NEWMAT::ColumnVector m_par;
    mutable std::vector<double> m_cfhist;
    mutable std::vector<NEWMAT::ColumnVector> m_parhist;
};

inline bool zero_cf_diff_conv(double cfo, double cfn, double cftol)
{
    return 2.0 * std::fabs(cfo - cfn) <= cftol * (std::fabs(cfo) + std::fabs(cfn) + 2.0e-16);
}

inline NonlinOut levmar(const NonlinParam &p, const NonlinCF &cfo)
{
    p.SetCF(cfo.cf(p.Par()));
    bool success = true;
    double olambda = 0.0;
    NEWMAT::ColumnVector g;
    boost::shared_ptr<BFMatrix> H;
    while (p.NextIter(success))
    {
        if (success)
        {
            g = cfo.grad(p.Par());
            H = cfo.hess(p.Par(), H);
        }
        for (int i = 1; i <= p.NPar(); i++)
        {
            if (p.GaussNewtonType() == LM_LM)
                H->Set(i, i, ((1.0 + p.Lambda()) / (1.0 + olambda)) * H->Peek(i, i));
            else
                H->Set(i, i, H->Peek(i, i) + p.Lambda() - olambda);
        }
        NEWMAT::ColumnVector step;
        double ncf = 0.0;
        bool inv_fail = false;
        try
        {
            step = -H->SolveForx(g, SYM_POSDEF, p.EquationSolverTol(), p.EquationSolverMaxIter());
            ncf = cfo.cf(p.Par() + step);
        }
        catch (...)
        {
            inv_fail = true;
        }
        if (!inv_fail && (success = (ncf < p.CF())))
        {
            olambda = 0.0;
            p.SetPar(p.Par() + step);
            p.SetLambda(p.Lambda() / 10.0);
            if (zero_cf_diff_conv(p.CF(), ncf, p.FractionalCFTolerance()))
            {
                p.SetCF(ncf);
                p.SetStatus(NL_CFCONV);
                return p.Status();
            }
            p.SetCF(ncf);
        }
        else
        {
            success = false;
            olambda = p.Lambda();
            p.SetLambda(10.0 * p.Lambda());
            if (p.Lambda() > p.LambdaConvergenceCriterion())
            {
                p.SetStatus(NL_LCONV);
                return p.Status();
            }
        }
    }
    p.SetStatus(NL_MAXITER);
    return p.Status();
}

inline NonlinOut nonlin(const NonlinParam &p, const NonlinCF &cfo)
{
    /* fabber only ever asks for NL_LM (inference_nlls.cc:135) */
    return levmar(p, cfo);
}
} // namespace MISCMATHS
#endif
