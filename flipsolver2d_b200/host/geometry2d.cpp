#include "geometry2d.h"

// geometry2d.cpp:54-73 restated: distance to the nearest edge, sign flipped once per edge whose
// three crossing conditions agree (all true or all false).
float Geometry2d::signedDistance(Vec3 point) const
{
    const size_t n = m_verts.size();
    const Vec3 d0 = point - m_verts[0];
    float best = d0.dot(d0);
    float sign = 1.0f;
    size_t prev = n - 1;
    for (size_t cur = 0; cur < n; prev = cur, cur++)
    {
        const Vec3 edge = m_verts[prev] - m_verts[cur];
        const Vec3 rel = point - m_verts[cur];
        const float t = std::clamp(rel.dot(edge) / edge.dot(edge), 0.0f, 1.0f);
        const Vec3 off = rel - edge * t;
        best = std::min(best, off.dot(off));
        const bool above = point.y() >= m_verts[cur].y();
        const bool below = point.y() < m_verts[prev].y();
        const bool side = edge.x() * rel.y() > edge.y() * rel.x();
        if ((above && below && side) || (!above && !below && !side)) sign *= -1.0f;
    }
    return sign * std::sqrt(best);
}
