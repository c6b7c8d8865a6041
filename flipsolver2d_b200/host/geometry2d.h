// Host mirror of FlipSolver2dLib/geometry2d.h: Vec3 (three floats) and the polygon type whose signed
// distance rasterises scene objects at frame 0 (geometry2d.cpp:54-73). Same names and semantics as the
// reference so callers (scene reader, GUI-style accessors) compile unchanged; arithmetic is written in
// plain float expressions and this library is compiled with -ffp-contract=off, i.e. it reproduces the
// reference built without -ffast-math.
#ifndef FS2D_HOST_GEOMETRY2D_H
#define FS2D_HOST_GEOMETRY2D_H

#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>

class Vec3
{
public:
    Vec3(float x = 0.0f, float y = 0.0f, float z = 0.0f) : m_x(x), m_y(y), m_z(z) {}
    Vec3(std::pair<float, float> &p) : m_x(p.first), m_y(p.second), m_z(0.f) {}

    float &x() { return m_x; }
    float &y() { return m_y; }
    float &z() { return m_z; }
    const float &x() const { return m_x; }
    const float &y() const { return m_y; }
    const float &z() const { return m_z; }

    float dot(Vec3 o) const { return m_x * o.m_x + m_y * o.m_y + m_z * o.m_z; }
    float distFromZero() const { return std::sqrt(m_x * m_x + m_y * m_y + m_z * m_z); }

    Vec3 normalized() const
    {
        const float len = distFromZero();
        if (std::abs(len) < 1e-6f) return Vec3();
        return Vec3(m_x / len, m_y / len, m_z / len);
    }

private:
    float m_x, m_y, m_z;
};

inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x() - b.x(), a.y() - b.y(), a.z() - b.z()); }
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x() + b.x(), a.y() + b.y(), a.z() + b.z()); }
inline Vec3 operator*(Vec3 a, float s) { return Vec3(a.x() * s, a.y() * s, a.z() * s); }
inline Vec3 operator*(float s, Vec3 a) { return Vec3(a.x() * s, a.y() * s, a.z() * s); }
inline Vec3 operator/(Vec3 a, float s) { return Vec3(a.x() / s, a.y() / s, a.z() / s); }

class Geometry2d
{
public:
    Geometry2d() = default;
    explicit Geometry2d(std::vector<Vec3> &verts) : m_verts(verts) {}

    void addVertex(Vec3 v) { m_verts.push_back(v); }
    std::vector<Vec3> verts() { return m_verts; }
    int vertextCount() { return static_cast<int>(m_verts.size()); }

    // Even-odd signed distance to the closed polygon (negative inside).
    float signedDistance(Vec3 point) const;
    float signedDistance(float x, float y) const { return signedDistance(Vec3(x, y)); }

private:
    std::vector<Vec3> m_verts;
};

#endif
