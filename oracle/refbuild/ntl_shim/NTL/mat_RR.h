// NTL/mat_RR.h stand-in — see RR.h in this directory (TEST INFRASTRUCTURE ONLY).
// The reference's Remez exchange (M/source/bootstrapping/common/Remez.cpp:176-216) builds a (deg+2)^2 system,
// transposes it, inverts it and multiplies a row vector by the inverse; that is all that is provided here.
#pragma once
#include "RR.h"
#include <stdexcept>
#include <vector>

namespace NTL
{
    class vec_RR
    {
    public:
        std::vector<RR> v;
        void SetLength(long n)
        {
            v.assign(size_t(n), RR(0));
        }
        long length() const
        {
            return long(v.size());
        }
        RR &operator[](long i)
        {
            return v[size_t(i)];
        }
        const RR &operator[](long i) const
        {
            return v[size_t(i)];
        }
    };

    class mat_RR
    {
    public:
        std::vector<vec_RR> rows;
        long r = 0, c = 0;
        void SetDims(long nr, long nc)
        {
            r = nr;
            c = nc;
            rows.assign(size_t(nr), vec_RR());
            for (auto &x : rows)
            {
                x.SetLength(nc);
            }
        }
        long NumRows() const
        {
            return r;
        }
        long NumCols() const
        {
            return c;
        }
        vec_RR &operator[](long i)
        {
            return rows[size_t(i)];
        }
        const vec_RR &operator[](long i) const
        {
            return rows[size_t(i)];
        }
    };

    inline void transpose(mat_RR &x, const mat_RR &a)
    {
        mat_RR t;
        t.SetDims(a.c, a.r);
        for (long i = 0; i < a.r; i++)
        {
            for (long j = 0; j < a.c; j++)
            {
                t[j][i] = a[i][j];
            }
        }
        x = t;
    }

    // x = a^{-1}, d = det(a): Gauss-Jordan with partial pivoting in RR
    inline void inv(RR &d, mat_RR &x, const mat_RR &a)
    {
        long n = a.r;
        if (a.c != n)
        {
            throw std::invalid_argument("inv: not square");
        }
        mat_RR w = a, id;
        id.SetDims(n, n);
        for (long i = 0; i < n; i++)
        {
            id[i][i] = RR(1);
        }
        d = RR(1);
        for (long col = 0; col < n; col++)
        {
            long piv = col;
            for (long i = col + 1; i < n; i++)
            {
                if (abs(w[i][col]) > abs(w[piv][col]))
                {
                    piv = i;
                }
            }
            if (w[piv][col] == RR(0))
            {
                d = RR(0);
                return;
            }
            if (piv != col)
            {
                std::swap(w.rows[size_t(piv)], w.rows[size_t(col)]);
                std::swap(id.rows[size_t(piv)], id.rows[size_t(col)]);
                d = -d;
            }
            d *= w[col][col];
            RR ip = RR(1) / w[col][col];
            for (long j = 0; j < n; j++)
            {
                w[col][j] *= ip;
                id[col][j] *= ip;
            }
            for (long i = 0; i < n; i++)
            {
                if (i == col || w[i][col] == RR(0))
                {
                    continue;
                }
                RR f = w[i][col];
                for (long j = 0; j < n; j++)
                {
                    w[i][j] -= f * w[col][j];
                    id[i][j] -= f * id[col][j];
                }
            }
        }
        x = id;
    }

    inline vec_RR operator*(const vec_RR &v, const mat_RR &m)
    {
        vec_RR out;
        out.SetLength(m.c);
        for (long j = 0; j < m.c; j++)
        {
            RR s(0);
            for (long i = 0; i < m.r; i++)
            {
                s += v[i] * m[i][j];
            }
            out[j] = s;
        }
        return out;
    }
} // namespace NTL
