"""woo.dem of the reference, restricted to the hot path of SURVEY.md §8a, plus GpuDemLoop."""
from .._woo import (DemField, DemData, Particle, ParticleContainer, ContactContainer, Contact, L6Geom, FrictPhys, Shape, Sphere, Wall, Facet,  # noqa: F401
                    Material, ElastMat, FrictMat, HertzMat, Leapfrog, InsertionSortCollider, Bo1_Sphere_Aabb, Bo1_Facet_Aabb, Bo1_Wall_Aabb,
                    ContactLoop, Cg2_Sphere_Sphere_L6Geom, Cg2_Wall_Sphere_L6Geom, Cg2_Facet_Sphere_L6Geom, Cp2_FrictMat_FrictPhys,
                    Cp2_HertzMat_HertzPhys, Law2_L6Geom_FrictPhys_IdealElPl, Law2_L6Geom_HertzPhys_DMT, DynDt, GpuDemLoop, ClumpData, Impose, HarmonicOscillation, AlignedHarmonicOscillations, ConstantForce, RadialForce,
                    Local6Dofs, VelocityAndReadForce, ReadForce, VariableVelocity3d, VariableAlignedRotation, CombinedImpose, CircularOrbit, StableCircularOrbit, InterpolatedMotion)
