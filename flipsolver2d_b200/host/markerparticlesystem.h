// Host mirror of FlipSolver2dLib/markerparticlesystem.h. On the device the particles are one SoA sorted
// by cell; the reference's 3x3-cell ParticleBin view (read by Liquid2dRender/fluidrenderer.cpp:934-986
// through markerParticles().bins().data()) is rebuilt from a download when somebody asks for it.
#ifndef FS2D_HOST_MARKERPARTICLESYSTEM_H
#define FS2D_HOST_MARKERPARTICLESYSTEM_H

#include <vector>

#include "grid2d.h"

class ParticleBin
{
public:
    size_t size() const { return m_positions.size(); }
    std::vector<Vec3> &positions() { return m_positions; }
    std::vector<Vec3> &velocities() { return m_velocities; }
    const Vec3 &particlePosition(size_t k) const { return m_positions[k]; }
    const Vec3 &particleVelocity(size_t k) const { return m_velocities[k]; }
    template <class T> std::vector<T> &particleProperties(size_t column) { return m_properties.at(column); }
    std::vector<std::vector<float>> &properties() { return m_properties; }
    void clear()
    {
        m_positions.clear();
        m_velocities.clear();
        for (auto &c : m_properties) c.clear();
    }

private:
    std::vector<Vec3> m_positions, m_velocities;
    std::vector<std::vector<float>> m_properties;
};

class MarkerParticleSystem
{
public:
    MarkerParticleSystem(size_t gridSizeI, size_t gridSizeJ, size_t binSize)
        : m_binSize(binSize), m_gridIndexer(gridSizeI, gridSizeJ),
          m_bins(gridSizeI / binSize + (gridSizeI % binSize != 0), gridSizeJ / binSize + (gridSizeJ % binSize != 0))
    {
    }
    Grid2d<ParticleBin> &bins() { return m_bins; }
    size_t particleCount() const { return m_count; }
    // markerparticlesystem.cpp:97-102,146-159
    Index2d binIdxForIdx(ssize_t i, ssize_t j) const { return Index2d(i / static_cast<ssize_t>(m_binSize), j / static_cast<ssize_t>(m_binSize)); }
    ssize_t gridToBinIdx(ssize_t i, ssize_t j) const { return m_bins.linearIndex(binIdxForIdx(i, j)); }
    ssize_t gridToBinIdx(Vec3 pos) const { return gridToBinIdx(static_cast<ssize_t>(pos.x()), static_cast<ssize_t>(pos.y())); }
    template <class T> size_t addParticleProperty() { return m_propertyCount++; }
    size_t propertyCount() const { return m_propertyCount; }

    // Rebuild the bin view from SoA arrays (pos/vel: 2 floats per particle, props: [column][count]).
    void assign(size_t count, const float *pos, const float *vel, const float *props)
    {
        for (ParticleBin &b : m_bins.data())
        {
            b.properties().resize(m_propertyCount);
            b.clear();
        }
        for (size_t p = 0; p < count; p++)
        {
            const Vec3 x(pos[2 * p], pos[2 * p + 1]);
            ParticleBin &b = m_bins.data()[gridToBinIdx(x)];
            b.positions().push_back(x);
            b.velocities().push_back(Vec3(vel[2 * p], vel[2 * p + 1]));
            for (size_t c = 0; c < m_propertyCount; c++) b.properties()[c].push_back(props[c * count + p]);
        }
        m_count = count;
    }
    void setCount(size_t n) { m_count = n; }

private:
    size_t m_binSize;
    LinearIndexable2d m_gridIndexer;
    Grid2d<ParticleBin> m_bins;
    size_t m_propertyCount = 0;
    size_t m_count = 0;
};

#endif
