"""Kinematics phase-space sampler (reference: `kinematics/pipeline.py`).

Same ``KinematicsPipeline`` constructor, validation, errors, ``run()`` result layout and
``run_kinematics_pipeline`` file layout as the reference.  The difference is ``run_batch(n)``:
all samples are drawn and all two-body boosts are done on arrays, with the reference's
resample-until-allowed rule applied to the rejected subset only, because the reference's
one-event-at-a-time loop cannot feed a GPU that simulates 10^5-10^6 events per second.
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np
from numpy.random import default_rng

from .angle import PolarDistribution
from .excitation import ExcitationDistribution
from .reaction import Decay, Reaction

CHUNK_SIZE: int = 1_000_000


@dataclass
class KinematicsTargetMaterial:
    """Gas target plus vertex sampling ranges (`pipeline.py:16-36`): z_range in m, rho_sigma in m."""

    material: object
    z_range: tuple[float, float]
    rho_sigma: float


@dataclass
class Sample:
    """Sampled parameters of one event (`pipeline.py:39-70`)."""

    beam_energy: float
    reaction_excitation: float
    reaction_theta: float
    reaction_phi: float
    vertex: np.ndarray
    decay_excitations: list[float]
    decay_thetas: list[float]
    decay_phis: list[float]


class PipelineError(Exception):
    """Pipeline error class"""


def _draw(dist, rng, n):
    if hasattr(dist, "sample_n"):
        return np.asarray(dist.sample_n(rng, n), dtype=np.float64)
    return np.array([dist.sample(rng) for _ in range(n)], dtype=np.float64)


class KinematicsPipeline:
    """Reaction followed by any number of sequential decays (`pipeline.py:79-426`)."""

    def __init__(
        self,
        steps: list[Reaction | Decay],
        excitations: list[ExcitationDistribution],
        polar_dists: list[PolarDistribution],
        beam_energy: float,
        target_material: KinematicsTargetMaterial | None = None,
        event_sample_limit: int = 1000,
    ):
        if len(steps) == 0:
            raise PipelineError("Pipeline must have at least one step (a Reaction)!")
        elif len(steps) != len(excitations):
            raise PipelineError(
                f"Pipeline must have the same number of steps (given {len(steps)}) and excitations (given {len(excitations)}!"
            )
        elif len(steps) != len(polar_dists):
            raise PipelineError(
                f"Pipeline must have the same number of steps (given {len(steps)}) and polar angle distributions (given {len(polar_dists)})!"
            )
        elif not isinstance(steps[0], Reaction):
            raise PipelineError("The first element in the pipeline must be a Reaction!")

        self.reaction: Reaction = steps[0]
        self.decays: list[Decay] = []
        self.excitations = excitations
        self.polar_dists = polar_dists
        self.rng = default_rng()
        self.event_sample_limit = event_sample_limit

        for idx in range(1, len(steps)):
            cur_step = steps[idx]
            if not isinstance(cur_step, Decay):
                raise PipelineError("All elements in the pipeline after the first element must be Decay!")
            prev_step = steps[idx - 1]
            produced = prev_step.residual if isinstance(prev_step, Reaction) else prev_step.residual_2
            if produced.isotopic_symbol != cur_step.parent.isotopic_symbol:
                which = "residual" if isinstance(prev_step, Reaction) else "residual_2"
                raise PipelineError(
                    f"Broken step in pipeline! Step {idx - 1} {which} does not match Step {idx} parent!"
                )
            self.decays.append(cur_step)

        returned_nuclei = 4 + (len(steps) - 1) * 2
        self.result = np.empty((returned_nuclei, 4), dtype=float)
        self.beam_energy = beam_energy
        self.target_material = target_material

    def __str__(self) -> str:
        chain = f"{self.reaction}"
        for decay in self.decays:
            chain += f", {str(decay)}"
        return chain

    def seed(self, seed: int) -> "KinematicsPipeline":
        """Make the sampling reproducible (the reference has no such hook, `pipeline.py:152`)."""
        self.rng = default_rng(seed)
        return self

    def check_excitations_allowed(self, projectile_energy: float, excitations: list[float]) -> bool:
        """Energy balance of the whole chain (`pipeline.py:200-230`)."""
        q_value = (
            (self.reaction.projectile.mass + projectile_energy)
            + self.reaction.target.mass
            - (self.reaction.ejectile.mass + self.reaction.residual.mass + excitations[0])
        )
        for idx, decay in enumerate(self.decays):
            q_value += -1.0 * (decay.residual_1.mass + decay.residual_2.mass + excitations[idx + 1])
        return q_value >= 0.0

    # ------------------------------------------------------------------------------- sampling
    def sample_batch(self, n: int) -> dict[str, np.ndarray]:
        """``n`` independent samples of every pipeline parameter (`pipeline.py:232-283`)."""
        rng = self.rng
        energy = np.full(n, float(self.beam_energy))
        vertex = np.zeros((n, 3))
        if self.target_material is not None:
            tm = self.target_material
            rho = np.abs(rng.normal(0.0, tm.rho_sigma, size=n))
            theta = rng.uniform(0.0, 2.0 * np.pi, size=n)
            vertex[:, 0] = rho * np.cos(theta)
            vertex[:, 1] = rho * np.sin(theta)
            vertex[:, 2] = rng.uniform(tm.z_range[0], tm.z_range[1], size=n)
            energy = energy - np.asarray(
                tm.material.get_energy_loss(self.reaction.projectile, float(self.beam_energy), vertex[:, 2])
            )
        k = len(self.excitations)
        return dict(
            beam_energy=energy,
            vertex=vertex,
            excitations=np.stack([_draw(d, rng, n) for d in self.excitations], axis=1),
            thetas=np.stack([_draw(d, rng, n) for d in self.polar_dists], axis=1),
            phis=rng.uniform(0.0, 2.0 * np.pi, size=(n, k)),
        )

    def sample(self) -> Sample:
        s = self.sample_batch(1)
        return Sample(
            beam_energy=float(s["beam_energy"][0]),
            reaction_excitation=float(s["excitations"][0, 0]),
            reaction_theta=float(s["thetas"][0, 0]),
            reaction_phi=float(s["phis"][0, 0]),
            vertex=s["vertex"][0],
            decay_excitations=[float(v) for v in s["excitations"][0, 1:]],
            decay_thetas=[float(v) for v in s["thetas"][0, 1:]],
            decay_phis=[float(v) for v in s["phis"][0, 1:]],
        )

    def _evaluate(self, s: dict[str, np.ndarray]) -> tuple[np.ndarray, np.ndarray]:
        """Kinematics of every sample and the mask of energetically allowed ones (`pipeline.py:320-386`)."""
        n = len(s["beam_energy"])
        out = np.zeros((n, len(self.result), 4))
        ok = np.asarray(self.reaction.is_excitation_allowed(s["beam_energy"], s["excitations"][:, 0])).reshape(n)
        safe_energy = np.where(ok, s["beam_energy"], self.beam_energy)
        with np.errstate(invalid="ignore", divide="ignore"):
            out[:, :4] = self.reaction.calculate_batch(safe_energy, s["thetas"][:, 0], s["phis"][:, 0], s["excitations"][:, 0])
            parent = out[:, 3]
            for idx, decay in enumerate(self.decays):
                ok &= np.asarray(decay.is_excitation_allowed(parent, s["excitations"][:, idx + 1])).reshape(n)
                r1, r2 = decay.calculate_batch(parent, s["thetas"][:, idx + 1], s["phis"][:, idx + 1], s["excitations"][:, idx + 1])
                out[:, 4 + 2 * idx] = r1
                out[:, 5 + 2 * idx] = r2
                parent = r2
        ok &= np.all(np.isfinite(out.reshape(n, -1)), axis=1)
        return out, ok

    def run_batch(self, n_events: int) -> tuple[np.ndarray, np.ndarray]:
        """``(vertices [n, 3], momenta [n, K, 4])`` with the resample-until-allowed rule of `run`."""
        vertices = np.zeros((n_events, 3))
        momenta = np.zeros((n_events, len(self.result), 4))
        todo = np.arange(n_events)
        for _ in range(self.event_sample_limit):
            if len(todo) == 0:
                break
            s = self.sample_batch(len(todo))
            out, ok = self._evaluate(s)
            vertices[todo[ok]] = s["vertex"][ok]
            momenta[todo[ok]] = out[ok]
            todo = todo[~ok]
        if len(todo):
            raise PipelineError(
                f"Reached Sampling Limit ({self.event_sample_limit} samples) for a single event! You may have defined an illegal reaction!"
            )
        return vertices, momenta

    def run(self) -> tuple[np.ndarray, np.ndarray]:
        """One event: ``(vertex [3], result [K, 4])`` rows px, py, pz, E (`pipeline.py:285-388`)."""
        vertices, momenta = self.run_batch(1)
        self.result[:] = momenta[0]
        return (vertices[0], self.result)

    def get_proton_numbers(self) -> np.ndarray:
        z = [self.reaction.target.Z, self.reaction.projectile.Z, self.reaction.ejectile.Z, self.reaction.residual.Z]
        for decay in self.decays:
            z += [decay.residual_1.Z, decay.residual_2.Z]
        return np.array(z, dtype=int)

    def get_mass_numbers(self) -> np.ndarray:
        a = [self.reaction.target.A, self.reaction.projectile.A, self.reaction.ejectile.A, self.reaction.residual.A]
        for decay in self.decays:
            a += [decay.residual_1.A, decay.residual_2.A]
        return np.array(a, dtype=int)


def save_kinematics_npz(path: Path, vertices: np.ndarray, momenta: np.ndarray, proton_numbers, mass_numbers) -> None:
    """Bulk kinematics file read by `detector.run_simulation` (no h5py needed)."""
    np.savez(
        path, data=np.asarray(momenta), vertices=np.asarray(vertices), proton_numbers=np.asarray(proton_numbers),
        mass_numbers=np.asarray(mass_numbers),
    )  # fmt: skip


def run_kinematics_pipeline(pipeline: KinematicsPipeline, n_events: int, output_path: Path, verbose: bool = True) -> None:
    """Sample ``n_events`` and write them (`pipeline.py:429-495`).

    ``*.npz`` -> one bulk file; anything else -> the reference's HDF5 layout (``/data`` attrs
    ``n_events, proton_numbers, mass_numbers, chunk_size, n_chunks``; groups ``chunk_i`` with attrs
    ``min_event, max_event``; datasets ``event_j`` [K, 4] with attrs ``vertex_x/y/z``), needs h5py.
    """
    output_path = Path(output_path)
    if verbose:
        print("------- AT-TPC Simulation Engine (B200) -------")
        print(f"Sampling kinematics from reaction: {pipeline}")
        print(f"Running for {n_events} samples.")
        print(f"Output will be written to {output_path}.")
    zs, as_ = pipeline.get_proton_numbers(), pipeline.get_mass_numbers()
    # sampled and written in slices of CHUNK_SIZE events, so that the temporaries of the sampler stay bounded
    # (the reference streams event by event); the .npz holds the whole run: 8 (4 K + 3) bytes per event
    if output_path.suffix == ".npz":
        momenta = np.empty((n_events, len(zs), 4), dtype=np.float64)
        vertices = np.empty((n_events, 3), dtype=np.float64)
        for lo in range(0, n_events, CHUNK_SIZE):
            hi = min(lo + CHUNK_SIZE, n_events)
            vertices[lo:hi], momenta[lo:hi] = pipeline.run_batch(hi - lo)
        save_kinematics_npz(output_path, vertices, momenta, zs, as_)
    else:
        try:
            import h5py
        except ImportError as exc:
            raise ImportError("writing HDF5 kinematics needs h5py; give an .npz path instead") from exc
        with h5py.File(output_path, "w") as output_file:
            data_group = output_file.create_group("data")
            data_group.attrs["n_events"] = n_events
            data_group.attrs["proton_numbers"] = zs
            data_group.attrs["mass_numbers"] = as_
            data_group.attrs["chunk_size"] = CHUNK_SIZE
            n_chunks = max(1, -(-n_events // CHUNK_SIZE))
            for chunk in range(n_chunks):
                lo, hi = chunk * CHUNK_SIZE, min((chunk + 1) * CHUNK_SIZE, n_events)
                chunk_group = data_group.create_group(f"chunk_{chunk}")
                chunk_group.attrs["min_event"] = lo
                chunk_group.attrs["max_event"] = hi - 1
                vertices, momenta = pipeline.run_batch(hi - lo) if hi > lo else (np.zeros((0, 3)), np.zeros((0, len(zs), 4)))
                for event in range(lo, hi):
                    data = chunk_group.create_dataset(f"event_{event}", data=momenta[event - lo])
                    data.attrs["vertex_x"] = vertices[event - lo, 0]
                    data.attrs["vertex_y"] = vertices[event - lo, 1]
                    data.attrs["vertex_z"] = vertices[event - lo, 2]
            data_group.attrs["n_chunks"] = n_chunks
    if verbose:
        print("Done.")
        print("----------------------------------------")
