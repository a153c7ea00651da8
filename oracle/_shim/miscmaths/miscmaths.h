/*
 * miscmaths.h - TEST INFRASTRUCTURE ONLY. Stand-in for the handful of FSL MISCMATHS functions the
 * reference's VB path calls (FSL is not vendored in the reference tree and not installed here); see
 * armawrap/newmat.h in this directory for why it exists.
 *   digamma             `float digamma(const float)` - FSL's source is not available; restated as
 *                       Bernardo's algorithm AS 103 in single precision (same restatement as the oracle)
 *   read_vest / read_ascii_matrix   design-matrix text files (tools.cc:27-40)
 *   sign
 */
#ifndef FABBER_SHIM_MISCMATHS_H
#define FABBER_SHIM_MISCMATHS_H

#include <cmath>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "armawrap/newmat.h"

namespace MISCMATHS
{
inline float digamma(const float x)
{
    const float s = 1e-5f, c = 8.5f, s3 = 8.333333333e-2f, s4 = 8.333333333e-3f, s5 = 3.968253968e-3f,
                d1 = -0.5772156649f;
    float y = x;
    float dg = 0.0f;
    if (y <= s)
        return d1 - 1.0f / y;
    while (y < c)
    {
        dg = dg - 1.0f / y;
        y = y + 1.0f;
    }
    float r = 1.0f / y;
    dg = (float)((double)dg + (double)(float)std::log((double)y) - 0.5 * (double)r);
    r = r * r;
    dg = dg - r * (s3 - r * (s4 - r * s5));
    return dg;
}

template <class T> inline int sign(const T &x) { return x > 0 ? 1 : (x < 0 ? -1 : 0); }

inline NEWMAT::Matrix read_rows(std::istream &in, const std::string &what)
{
    std::vector<std::vector<double> > rows;
    std::string line;
    while (std::getline(in, line))
    {
        size_t first = line.find_first_not_of(" \t\r");
        if (first == std::string::npos)
            continue;
        if (line[first] == '#' || line[first] == '%')
            continue;
        std::istringstream s(line);
        std::vector<double> row;
        double x;
        while (s >> x)
            row.push_back(x);
        if (row.empty())
            throw NEWMAT::Exception("non-numeric line in matrix file " + what);
        if (!rows.empty() && row.size() != rows[0].size())
            throw NEWMAT::Exception("ragged matrix file " + what);
        rows.push_back(row);
    }
    if (rows.empty())
        throw NEWMAT::Exception("empty matrix file " + what);
    NEWMAT::Matrix m((int)rows.size(), (int)rows[0].size());
    for (size_t i = 0; i < rows.size(); i++)
        for (size_t j = 0; j < rows[i].size(); j++)
            m((int)i + 1, (int)j + 1) = rows[i][j];
    return m;
}

inline NEWMAT::Matrix read_vest(const std::string &filename)
{
    std::ifstream in(filename.c_str());
    if (!in)
        throw NEWMAT::Exception("could not open " + filename);
    std::string line;
    bool found = false;
    while (std::getline(in, line))
        if (line.compare(0, 7, "/Matrix") == 0)
        {
            found = true;
            break;
        }
    if (!found)
        throw NEWMAT::Exception("not a VEST file: " + filename);
    return read_rows(in, filename);
}

inline NEWMAT::Matrix read_ascii_matrix(const std::string &filename)
{
    std::ifstream in(filename.c_str());
    if (!in)
        throw NEWMAT::Exception("could not open " + filename);
    return read_rows(in, filename);
}
} // namespace MISCMATHS

#endif
