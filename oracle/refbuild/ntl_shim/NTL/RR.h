// NTL/RR.h stand-in — TEST INFRASTRUCTURE ONLY (oracle/_ref), never part of the product.
//
// The reference's bootstrapping sources (M/source/bootstrapping/{Bootstrapper,ModularReducer}.cpp,
// common/{Polynomial,Remez,func,Point}.cpp) use NTL's arbitrary-precision float `RR` (Polynomial.h:3,
// Remez.h:3, func.h:3-4).  NTL and its headers are not in this image; libmpfr.so.6 (a run-time
// dependency of gcc) is.  This header provides the part of NTL's RR / ZZ interface those files use on top
// of MPFR, with the handful of MPFR prototypes declared by hand (MPFR 4 ABI: __mpfr_struct, mpfr_rnd_t).
// Every value carries RR_SHIM_PREC bits, at least RemezParam::RR_prec = 1000 (RemezParam.h:13), so the
// reference's Remez exchange and its power <-> Chebyshev basis conversions run at the precision the
// reference asks for (a 113-bit __float128 loses the polynomial in Polynomial::cheb_to_power).
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

extern "C"
{
    typedef struct
    {
        long _mpfr_prec;
        int _mpfr_sign;
        long _mpfr_exp;
        unsigned long *_mpfr_d;
    } ntlshim_mpfr_struct;
    typedef ntlshim_mpfr_struct *ntlshim_mpfr_ptr;
    typedef const ntlshim_mpfr_struct *ntlshim_mpfr_srcptr;
    void mpfr_init2(ntlshim_mpfr_ptr, long);
    void mpfr_clear(ntlshim_mpfr_ptr);
    int mpfr_set(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr, int);
    int mpfr_set_d(ntlshim_mpfr_ptr, double, int);
    int mpfr_set_si(ntlshim_mpfr_ptr, long, int);
    int mpfr_set_str(ntlshim_mpfr_ptr, const char *, int, int);
    double mpfr_get_d(ntlshim_mpfr_srcptr, int);
    long mpfr_get_si(ntlshim_mpfr_srcptr, int);
    int mpfr_add(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr, ntlshim_mpfr_srcptr, int);
    int mpfr_sub(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr, ntlshim_mpfr_srcptr, int);
    int mpfr_mul(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr, ntlshim_mpfr_srcptr, int);
    int mpfr_div(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr, ntlshim_mpfr_srcptr, int);
    int mpfr_neg(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr, int);
    int mpfr_abs(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr, int);
    int mpfr_sqrt(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr, int);
    int mpfr_cos(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr, int);
    int mpfr_sin(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr, int);
    int mpfr_pow(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr, ntlshim_mpfr_srcptr, int);
    int mpfr_round(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr);
    int mpfr_floor(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr);
    int mpfr_const_pi(ntlshim_mpfr_ptr, int);
    int mpfr_mul_2si(ntlshim_mpfr_ptr, ntlshim_mpfr_srcptr, long, int);
    int mpfr_cmp(ntlshim_mpfr_srcptr, ntlshim_mpfr_srcptr);
    int mpfr_asprintf(char **, const char *, ...);
    void mpfr_free_str(char *);
}

#ifndef RR_SHIM_PREC
#define RR_SHIM_PREC 1024
#endif

namespace NTL
{
    class ZZ
    {
    public:
        long v = 0;
        ZZ() {}
        ZZ(long x) : v(x) {}
        ZZ &operator=(long x)
        {
            v = x;
            return *this;
        }
    };
    inline long operator%(const ZZ &a, long m)
    {
        long r = a.v % m;
        return r < 0 ? r + m : r;
    }

    class RR
    {
    public:
        mutable ntlshim_mpfr_struct m;
        RR()
        {
            mpfr_init2(&m, RR_SHIM_PREC);
            mpfr_set_si(&m, 0, 0);
        }
        RR(const RR &o)
        {
            mpfr_init2(&m, RR_SHIM_PREC);
            mpfr_set(&m, &o.m, 0);
        }
        RR(double d)
        {
            mpfr_init2(&m, RR_SHIM_PREC);
            mpfr_set_d(&m, d, 0);
        }
        RR(int i)
        {
            mpfr_init2(&m, RR_SHIM_PREC);
            mpfr_set_si(&m, i, 0);
        }
        RR(long i)
        {
            mpfr_init2(&m, RR_SHIM_PREC);
            mpfr_set_si(&m, i, 0);
        }
        RR(long long i)
        {
            mpfr_init2(&m, RR_SHIM_PREC);
            mpfr_set_si(&m, long(i), 0);
        }
        RR(unsigned long i)
        {
            mpfr_init2(&m, RR_SHIM_PREC);
            mpfr_set_d(&m, double(i), 0);
        }
        ~RR()
        {
            mpfr_clear(&m);
        }
        RR &operator=(const RR &o)
        {
            if (this != &o)
            {
                mpfr_set(&m, &o.m, 0);
            }
            return *this;
        }
        RR &operator+=(const RR &o)
        {
            mpfr_add(&m, &m, &o.m, 0);
            return *this;
        }
        RR &operator-=(const RR &o)
        {
            mpfr_sub(&m, &m, &o.m, 0);
            return *this;
        }
        RR &operator*=(const RR &o)
        {
            mpfr_mul(&m, &m, &o.m, 0);
            return *this;
        }
        RR &operator/=(const RR &o)
        {
            mpfr_div(&m, &m, &o.m, 0);
            return *this;
        }
        // NTL's knobs; this shim always computes with RR_SHIM_PREC bits (>= what the callers request)
        static void SetPrecision(long)
        {}
        static long precision()
        {
            return RR_SHIM_PREC;
        }
        static long &output_digits()
        {
            static long d = 10;
            return d;
        }
        static void SetOutputPrecision(long d)
        {
            output_digits() = d;
        }
    };

#define NTLSHIM_BINOP(op, fn)                                                                                          \
    inline RR operator op(const RR &a, const RR &b)                                                                   \
    {                                                                                                                  \
        RR r;                                                                                                          \
        fn(&r.m, &a.m, &b.m, 0);                                                                                       \
        return r;                                                                                                      \
    }                                                                                                                  \
    inline RR operator op(const RR &a, double b)                                                                      \
    {                                                                                                                  \
        return a op RR(b);                                                                                             \
    }                                                                                                                  \
    inline RR operator op(double a, const RR &b)                                                                      \
    {                                                                                                                  \
        return RR(a) op b;                                                                                             \
    }                                                                                                                  \
    inline RR operator op(const RR &a, long b)                                                                        \
    {                                                                                                                  \
        return a op RR(b);                                                                                             \
    }                                                                                                                  \
    inline RR operator op(long a, const RR &b)                                                                        \
    {                                                                                                                  \
        return RR(a) op b;                                                                                             \
    }                                                                                                                  \
    inline RR operator op(const RR &a, int b)                                                                         \
    {                                                                                                                  \
        return a op RR(b);                                                                                             \
    }                                                                                                                  \
    inline RR operator op(int a, const RR &b)                                                                         \
    {                                                                                                                  \
        return RR(a) op b;                                                                                             \
    }
    NTLSHIM_BINOP(+, mpfr_add)
    NTLSHIM_BINOP(-, mpfr_sub)
    NTLSHIM_BINOP(*, mpfr_mul)
    NTLSHIM_BINOP(/, mpfr_div)
#undef NTLSHIM_BINOP

    inline RR operator-(const RR &a)
    {
        RR r;
        mpfr_neg(&r.m, &a.m, 0);
        return r;
    }

#define NTLSHIM_CMP(op)                                                                                                \
    inline bool operator op(const RR &a, const RR &b)                                                                 \
    {                                                                                                                  \
        return mpfr_cmp(&a.m, &b.m) op 0;                                                                              \
    }                                                                                                                  \
    inline bool operator op(const RR &a, double b)                                                                    \
    {                                                                                                                  \
        return a op RR(b);                                                                                             \
    }                                                                                                                  \
    inline bool operator op(double a, const RR &b)                                                                    \
    {                                                                                                                  \
        return RR(a) op b;                                                                                             \
    }                                                                                                                  \
    inline bool operator op(const RR &a, long b)                                                                      \
    {                                                                                                                  \
        return a op RR(b);                                                                                             \
    }                                                                                                                  \
    inline bool operator op(long a, const RR &b)                                                                      \
    {                                                                                                                  \
        return RR(a) op b;                                                                                             \
    }                                                                                                                  \
    inline bool operator op(const RR &a, int b)                                                                       \
    {                                                                                                                  \
        return a op RR(b);                                                                                             \
    }                                                                                                                  \
    inline bool operator op(int a, const RR &b)                                                                       \
    {                                                                                                                  \
        return RR(a) op b;                                                                                             \
    }
    NTLSHIM_CMP(<)
    NTLSHIM_CMP(>)
    NTLSHIM_CMP(<=)
    NTLSHIM_CMP(>=)
    NTLSHIM_CMP(==)
    NTLSHIM_CMP(!=)
#undef NTLSHIM_CMP

#define NTLSHIM_UNARY(name, fn)                                                                                        \
    inline RR name(const RR &a)                                                                                       \
    {                                                                                                                  \
        RR r;                                                                                                          \
        fn(&r.m, &a.m, 0);                                                                                             \
        return r;                                                                                                      \
    }
    NTLSHIM_UNARY(abs, mpfr_abs)
    NTLSHIM_UNARY(sqrt, mpfr_sqrt)
    NTLSHIM_UNARY(cos, mpfr_cos)
    NTLSHIM_UNARY(sin, mpfr_sin)
#undef NTLSHIM_UNARY

    inline RR round(const RR &a)
    {
        RR r;
        mpfr_round(&r.m, &a.m);
        return r;
    }
    inline RR floor(const RR &a)
    {
        RR r;
        mpfr_floor(&r.m, &a.m);
        return r;
    }
    inline ZZ RoundToZZ(const RR &a)
    {
        RR r = round(a);
        return ZZ(mpfr_get_si(&r.m, 0));
    }
    inline RR pow(const RR &a, const RR &b)
    {
        RR r;
        mpfr_pow(&r.m, &a.m, &b.m, 0);
        return r;
    }
    inline RR ComputePi_RR()
    {
        RR r;
        mpfr_const_pi(&r.m, 0);
        return r;
    }
    inline RR power2_RR(long e)
    {
        RR r(1);
        mpfr_mul_2si(&r.m, &r.m, e, 0);
        return r;
    }
    inline double to_double(const RR &a)
    {
        return mpfr_get_d(&a.m, 0);
    }
    inline RR to_RR(const RR &a)
    {
        return a;
    }
    inline RR to_RR(double a)
    {
        return RR(a);
    }
    inline RR to_RR(long a)
    {
        return RR(a);
    }
    inline RR to_RR(int a)
    {
        return RR(a);
    }

    inline std::ostream &operator<<(std::ostream &os, const RR &a)
    {
        char *s = nullptr;
        mpfr_asprintf(&s, "%.*Rg", int(RR::output_digits()), &a.m);
        if (s)
        {
            os << s;
            mpfr_free_str(s);
        }
        return os;
    }
    inline std::istream &operator>>(std::istream &is, RR &a)
    {
        std::string tok;
        is >> tok;
        if (is)
        {
            mpfr_set_str(&a.m, tok.c_str(), 10, 0);
        }
        return is;
    }
} // namespace NTL
