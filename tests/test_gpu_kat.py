"""The known-answer tests of tests/test_oracle_kat.py (reference formulas, ports of the reference's own tests, analytic cases) run a second time with the CUDA
engine standing in for the oracle: the same expectations must hold for the product path through the C-ABI."""
import pytest
import test_oracle_kat as K

pytestmark = pytest.mark.gpu


class GpuAsOracle:
    """woo_b200.engine.Engine behind the few methods of tests/_harness.Oracle that the KATs use."""

    def __init__(self, S):
        from woo_b200.engine import Engine
        self.E = Engine(S.params).load(S)

    def step(self, n=1): self.E.step(n)
    def contacts(self, real_only=False): return self.E.contacts(real_only)
    def spheres(self): return self.E.spheres()
    def num_contacts(self, real_only=False): return len(self.E.contacts(real_only)["ids"])
    def track_energy(self, on=True): self.E.track_energy(on)
    def energy(self): return self.E.energy()
    def stress(self): return self.E.stress()
    def kineticEnergy(self): return self.E.kinetic_energy()
    def cell_hsize(self): return self.E.cell_hsize()
    def pWaveDt(self): return self.E.crit_dt()
    def facets(self): return self.E.facets()
    def impose_state(self, ix): return self.E.impose_state(ix)
    @property
    def nFullRuns(self): return self.E.scalars()["nFullRuns"]
    @property
    def dt(self): return self.E.scalars()["dt"]
    @property
    def time(self): return self.E.scalars()["time"]


CASES = [(K.test_pwave_dt, ()), (K.test_cundall_kn_and_hertz_force, ()), (K.test_update_phys_kn_proportional_to_young, ()), (K.test_force_sign_ten_spheres, ()),
         (K.test_dyndt_two_sphere_analytic, ()), (K.test_leapfrog_damping_gravity_analytic, ()), (K.test_coulomb_limit_and_noslip, ()),
         (K.test_sphere_resting_on_wall_equilibrium, ()), (K.test_energy_conservation_elastic_collision, ()), (K.test_energy_free_fall_and_damping_like_the_reference_test, ()),
         (K.test_stress_two_spheres_across_the_cell_boundary, ()), (K.test_pbc_velgrad_changes_hsize, ()), (K.test_pbc_homothetic_resize_vel_and_pos, ()),
         (K.test_pbc_kinetic_energy_is_of_the_fluctuation_velocity, (K.sc.HOMO_VEL_2ND,)), (K.test_pbc_kinetic_energy_is_of_the_fluctuation_velocity, (K.sc.HOMO_POS,)),
         (K.test_facet_sphere_closest_point, ((.25, .25, .19), (.25, .25, 0.), "face")), (K.test_facet_sphere_closest_point, ((.5, -.1, .1), (.5, 0., 0.), "edge")),
         (K.test_facet_sphere_closest_point, ((1.1, -.1, .05), (1., 0., 0.), "vertex")), (K.test_hertz_collision_time, (1e-2,)),
         (K.test_hertz_restitution_coefficient, (.5,)), (K.test_collision_masks, (0b01, 0b10, False)), (K.test_collision_masks, (0b10, 0b10, False)),
         (K.test_collision_masks, (0b0101, 0b0011, True)), (K.test_collision_masks, (0b0111, 0b0011, False)), (K.test_collision_masks, (0b0101, 0b0111, True)),
         (K.test_impose_circular_orbits_keep_radius_and_turn_by_omega_t, ()), (K.test_impose_interpolated_motion_reaches_the_given_poses_at_the_given_times, ()),
         (K.test_facet_contact_through_the_periodic_boundary_equals_the_unwrapped_one, ())]


@pytest.mark.parametrize("kat,args", CASES, ids=[f"{f.__name__}{list(a) if a else ''}" for f, a in CASES])
def test_kat_with_the_cuda_engine(kat, args, monkeypatch):
    monkeypatch.setattr(K, "Oracle", GpuAsOracle)
    kat(*args)
