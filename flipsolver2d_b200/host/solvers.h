// FlipSolver2dLib/solvers.h
#ifndef FS2D_HOST_SOLVERS_H
#define FS2D_HOST_SOLVERS_H
#include "flipfiresolver.h"
#include "flipsmokesolver.h"
#include "flipsolver2d.h"
#include "nbflipsolver.h"
#endif
