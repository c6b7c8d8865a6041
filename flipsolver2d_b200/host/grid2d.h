// Host mirrors of the reference's dense grid types (FlipSolver2dLib/linearindexable2d.h, index2d.h,
// grid2d.h, materialgrid.h, sdfgrid.h, staggeredvelocitygrid.h). On the device every grid is a plain
// row-major array (idx = i*sizeJ + j, linearindexable2d.h:30-37); these classes hold the HOST copy that
// the solver's const accessors hand out, refreshed from the device on demand. Only the read API the
// reference's callers use (Liquid2dRender/fluidrenderer.cpp:497-673) plus what frame-0 seeding needs is
// mirrored.
#ifndef FS2D_HOST_GRID2D_H
#define FS2D_HOST_GRID2D_H

#include <sys/types.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <type_traits>
#include <vector>

#include "geometry2d.h"

struct Index2d
{
    Index2d(ssize_t i_ = 0, ssize_t j_ = 0) : i(i_), j(j_) {}
    ssize_t i, j;
};

struct Range
{
    Range(size_t s = 0, size_t e = 0) : start(s), end(e) {}
    size_t start, end;
    size_t size() const { return end - start; }
};

class LinearIndexable2d
{
public:
    LinearIndexable2d(size_t sizeI, size_t sizeJ) : m_sizeI(sizeI), m_sizeJ(sizeJ) {}
    ssize_t sizeI() const { return m_sizeI; }
    ssize_t sizeJ() const { return m_sizeJ; }
    ssize_t linearIndex(ssize_t i, ssize_t j) const { return inBounds(i, j) ? i * m_sizeJ + j : -1; }
    ssize_t linearIndex(Index2d idx) const { return linearIndex(idx.i, idx.j); }
    Index2d index2d(ssize_t lin) const
    {
        const ssize_t i = lin / m_sizeJ;
        return Index2d(i, lin - i * m_sizeJ);
    }
    bool inBounds(ssize_t i, ssize_t j) const { return i >= 0 && i < m_sizeI && j >= 0 && j < m_sizeJ; }
    bool inBounds(Index2d idx) const { return inBounds(idx.i, idx.j); }
    bool inBounds(ssize_t lin) const { return lin >= 0 && lin < m_sizeI * m_sizeJ; }
    size_t linearSize() const { return static_cast<size_t>(m_sizeI * m_sizeJ); }
    ssize_t linearIdxOfOffset(ssize_t lin, ssize_t di, ssize_t dj) const { return lin + di * m_sizeJ + dj; }

protected:
    ssize_t m_sizeI, m_sizeJ;
};

enum OOBStrategy : char { OOB_EXTEND, OOB_CONST, OOB_ERROR };

namespace simmath
{
inline float frac(float v) { return v - static_cast<long>(v); }
inline int integr(float v) { return static_cast<int>(std::floor(v)); }
inline float lerp(float a, float b, float f) { return (a * (1.0f - f)) + (b * f); }
}  // namespace simmath

template <class T> class Grid2d : public LinearIndexable2d
{
public:
    // bool grids are stored one byte per flag (the device layout), not as std::vector<bool>
    using Stored = typename std::conditional<std::is_same<T, bool>::value, uint8_t, T>::type;

    Grid2d(size_t sizeI, size_t sizeJ, T initValue = T(), OOBStrategy oobStrat = OOB_ERROR, T oobVal = T(),
           Vec3 gridOffset = Vec3(0.f, 0.f))
        : LinearIndexable2d(sizeI, sizeJ), m_oobStrat(oobStrat), m_data(sizeI * sizeJ, static_cast<Stored>(initValue)),
          m_oobConst(oobVal), m_gridOffset(gridOffset)
    {
    }

    OOBStrategy &oobStrat() { return m_oobStrat; }
    Vec3 &gridOffset() { return m_gridOffset; }

    Stored &at(ssize_t i, ssize_t j) { return m_data[linearIndex(i, j)]; }
    Stored &at(Index2d idx) { return m_data[linearIndex(idx)]; }
    const Stored &at(ssize_t i, ssize_t j) const { return m_data[linearIndex(i, j)]; }
    const Stored &at(Index2d idx) const { return m_data[linearIndex(idx)]; }
    void setAt(ssize_t i, ssize_t j, T v) { m_data[linearIndex(i, j)] = static_cast<Stored>(v); }

    // grid2d.h:122-143
    T getAt(ssize_t i, ssize_t j) const
    {
        if (m_oobStrat == OOB_EXTEND)
        {
            i = std::clamp<ssize_t>(i, 0, m_sizeI - 1);
            j = std::clamp<ssize_t>(j, 0, m_sizeJ - 1);
        }
        else if (m_oobStrat == OOB_CONST && !inBounds(i, j))
        {
            return m_oobConst;
        }
        return static_cast<T>(m_data[i * m_sizeJ + j]);
    }
    T getAt(Index2d idx) const { return getAt(idx.i, idx.j); }

    void fill(T v) { m_data.assign(m_data.size(), static_cast<Stored>(v)); }
    T oobVal() const { return m_oobConst; }
    std::vector<Stored> &data() { return m_data; }
    const std::vector<Stored> &data() const { return m_data; }

    // Grid2d::lerp (grid2d.h:187-216): cell-centred blend with |frac - 1/2| factors
    template <class U = T, typename std::enable_if<std::is_floating_point<U>::value>::type * = nullptr>
    T interpolateAt(float i, float j) const
    {
        i += m_gridOffset.x();
        j += m_gridOffset.y();
        i = std::clamp(i, 0.f, static_cast<float>(m_sizeI - 1));
        j = std::clamp(j, 0.f, static_cast<float>(m_sizeJ - 1));
        const ssize_t ci = simmath::integr(i), cj = simmath::integr(j);
        const float fi = simmath::frac(i), fj = simmath::frac(j);
        const ssize_t ni = fi >= 0.5f ? ci + 1 : ci - 1;
        const ssize_t nj = fj >= 0.5f ? cj + 1 : cj - 1;
        const float wi = fi < 0.5f ? 0.5f - fi : fi - 0.5f;
        const float wj = fj < 0.5f ? 0.5f - fj : fj - 0.5f;
        const float a = simmath::lerp(getAt(ci, cj), getAt(ni, cj), wi);
        const float b = simmath::lerp(getAt(ci, nj), getAt(ni, nj), wi);
        return simmath::lerp(a, b, wj);
    }
    template <class U = T, typename std::enable_if<std::is_floating_point<U>::value>::type * = nullptr>
    T interpolateAt(Vec3 p) const
    {
        return interpolateAt(p.x(), p.y());
    }
    template <class U = T, typename std::enable_if<std::is_floating_point<U>::value>::type * = nullptr>
    T lerpolateAt(float i, float j) const
    {
        return interpolateAt(i, j);
    }
    template <class U = T, typename std::enable_if<std::is_floating_point<U>::value>::type * = nullptr>
    T lerpolateAt(Vec3 p) const
    {
        return interpolateAt(p.x(), p.y());
    }

protected:
    OOBStrategy m_oobStrat;
    std::vector<Stored> m_data;
    T m_oobConst;
    Vec3 m_gridOffset;
};

// materialgrid.h:6-20
enum FluidMaterial : int8_t { FLUID = 0x40, SOURCE = 0x41, SOLID = 0x20, SINK = 0x12, EMPTY = 0x10 };

class MaterialGrid : public Grid2d<FluidMaterial>
{
public:
    // materialgrid.cpp:5-8: OOB_EXTEND, the oobMaterial argument is never looked up
    MaterialGrid(size_t sizeI, size_t sizeJ, FluidMaterial oobMaterial = SINK)
        : Grid2d<FluidMaterial>(sizeI, sizeJ, EMPTY, OOB_EXTEND, oobMaterial)
    {
    }
    bool isFluid(ssize_t i, ssize_t j) const { return (getAt(i, j) & 0x40) != 0; }
    bool isStrictFluid(ssize_t i, ssize_t j) const { return getAt(i, j) == FLUID; }
    bool isEmpty(ssize_t i, ssize_t j) const { return (getAt(i, j) & 0x10) != 0; }
    bool isSolid(ssize_t i, ssize_t j) const { return (getAt(i, j) & 0x20) != 0; }
    bool isSource(ssize_t i, ssize_t j) const { return getAt(i, j) == SOURCE; }
    bool isSink(ssize_t i, ssize_t j) const { return getAt(i, j) == SINK; }
    bool isFluid(Index2d x) const { return isFluid(x.i, x.j); }
    bool isEmpty(Index2d x) const { return isEmpty(x.i, x.j); }
    bool isSolid(Index2d x) const { return isSolid(x.i, x.j); }
    bool isSource(Index2d x) const { return isSource(x.i, x.j); }
    bool isSink(Index2d x) const { return isSink(x.i, x.j); }
};

class SdfGrid : public Grid2d<float>
{
public:
    SdfGrid(size_t sizeI, size_t sizeJ) : Grid2d<float>(sizeI, sizeJ, 0.f, OOB_EXTEND) {}
};

class StaggeredVelocityGrid : public LinearIndexable2d
{
public:
    // staggeredvelocitygrid.cpp:6-14
    StaggeredVelocityGrid(size_t sizeI, size_t sizeJ)
        : LinearIndexable2d(sizeI, sizeJ), m_u(sizeI + 1, sizeJ, 0.f, OOB_EXTEND, 0.f, Vec3(0.5f, 0.f)),
          m_v(sizeI, sizeJ + 1, 0.f, OOB_EXTEND, 0.f, Vec3(0.f, 0.5f)), m_uValid(sizeI + 1, sizeJ, false, OOB_CONST, true),
          m_vValid(sizeI, sizeJ + 1, false, OOB_CONST, true)
    {
    }
    Grid2d<float> &velocityGridU() { return m_u; }
    Grid2d<float> &velocityGridV() { return m_v; }
    const Grid2d<float> &velocityGridU() const { return m_u; }
    const Grid2d<float> &velocityGridV() const { return m_v; }
    Grid2d<bool> &uSampleValidityGrid() { return m_uValid; }
    Grid2d<bool> &vSampleValidityGrid() { return m_vValid; }
    float getU(ssize_t i, ssize_t j) const { return m_u.getAt(i, j); }
    float getV(ssize_t i, ssize_t j) const { return m_v.getAt(i, j); }
    float getU(Index2d x) const { return getU(x.i, x.j); }
    float getV(Index2d x) const { return getV(x.i, x.j); }
    Vec3 velocityAt(float i, float j) const { return Vec3(m_u.interpolateAt(i, j), m_v.interpolateAt(i, j)); }
    Vec3 velocityAt(Vec3 p) const { return velocityAt(p.x(), p.y()); }

private:
    Grid2d<float> m_u, m_v;
    Grid2d<bool> m_uValid, m_vValid;
};

#endif
