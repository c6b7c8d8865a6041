// Host mirrors of FlipSolver2dLib/emitter.h, obstacle.h and sink.h: property bags around a polygon.
#ifndef FS2D_HOST_SCENEOBJECTS_H
#define FS2D_HOST_SCENEOBJECTS_H

#include "geometry2d.h"

class Emitter
{
public:
    explicit Emitter(Geometry2d &geo) : m_geometry(geo) {}

    void setViscosity(float v) { m_viscosity = v; }
    float viscosity() const { return m_viscosity; }
    Geometry2d &geometry() { return m_geometry; }
    void setGeometry(Geometry2d &g) { m_geometry = g; }
    float temperature() const { return m_temperature; }
    void setTemperature(float v) { m_temperature = v; }
    float concentrartion() const { return m_concentration; }  // spelling as in emitter.h:21
    void setConcentrartion(float v) { m_concentration = v; }
    float divergence() const { return m_divergence; }
    void setDivergence(float v) { m_divergence = v; }
    float fuel() const { return m_fuel; }
    void setFuel(float v) { m_fuel = v; }
    Vec3 velocity() const { return m_velocity; }
    void setVelocity(Vec3 v) { m_velocity = v; }
    bool velocityTransfer() const { return m_transferVelocity; }
    void setVelocityTransfer(bool b) { m_transferVelocity = b; }

private:
    float m_viscosity = 0.f, m_temperature = 273.f, m_concentration = 1.f, m_divergence = 0.f, m_fuel = 1.f;
    bool m_transferVelocity = false;
    Vec3 m_velocity;
    Geometry2d m_geometry;
};

class Obstacle
{
public:
    Obstacle(float friction, Geometry2d &geo) : m_friction(friction), m_geometry(geo) {}
    float friction() { return m_friction; }
    void setFriction(float f) { m_friction = f; }
    Geometry2d &geometry() { return m_geometry; }
    void setGeometry(Geometry2d &g) { m_geometry = g; }

private:
    float m_friction;
    Geometry2d m_geometry;
};

class Sink
{
public:
    Sink(float divergence, Geometry2d &geo) : m_geo(geo), m_divergence(divergence) {}
    float divergence() const { return m_divergence; }
    void setDivergence(float d) { m_divergence = d; }
    Geometry2d &geo() { return m_geo; }
    void setGeo(const Geometry2d &g) { m_geo = g; }

private:
    Geometry2d m_geo;
    float m_divergence;
};

#endif
