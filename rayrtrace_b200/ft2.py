"""FT-2 interferometer geometry and the density-scan driver (the reference's ft2.py:1-71),
on top of the batched Beam.  Host-side O(beams) orchestration; every ray is traced on the GPU.

Only what sits on the path from a density grid to interferometer phases is mirrored
(``ants``, ``limiter``, ``phases``); the CX / TS chord helpers (ft2.py:74-90) never touch the
tracer and are out of scope.
"""
from __future__ import annotations

import numpy as np

from .beam import Beam, field_result, field_result_device
from .nfield import Nfield

r_dia = 7.85     # limiter radius, cm (ft2.py:6)
r_ant = 10       # antenna radius, cm (ft2.py:7)


def ants():
    """Seven antennas on the r = 10 cm circle, x from -6.9 to 6.9 cm (ft2.py:12-16)."""
    x = np.linspace(-6.9, 6.9, 7)
    return np.vstack((x, np.sqrt(r_ant ** 2 - x ** 2))).T


def limiter(r_dia=r_dia, theta=360):
    if isinstance(theta, (int, np.integer)):
        theta = np.linspace(0, 2 * np.pi, theta, endpoint=False)
    return r_dia * np.cos(theta), r_dia * np.sin(theta)


def _pair_phases(nfield, tx, rx, device=True):
    """One TX/RX antenna pair: trace the TX beam through ``nfield``, overlap its field on the
    receiver line with the RX antenna pattern (beam.py:447-472).  ``device=True``: the crossing of
    the receiver line is fused into the trace and the neighbour-ray lookup is one more launch --
    no trajectories kept, no Ray objects, and the O(rays) post-processing (pages, ray-tube amplitudes,
    overlap with the RX pattern, Gaussian fit) stays on the GPU (``field_result_device``);
    ``device="host-tail"``: the same two launches with the round-1 NumPy / SciPy post-processing;
    ``device=False``: the reference's flow on materialised rays (Ray.addpoint + interp1d on the host)."""
    beam_rx = Beam(x0=rx[0], y0=rx[1], kx0=0, ky0=-1)
    if device:
        beam_tx = Beam(x0=tx[0], y0=tx[1], kx0=0, ky0=1)
        gpu_tail = device != "host-tail"
        beam_tx.rays(n_field=nfield, keep_traj=False, receiver=(float(rx[1]), "y"), host_results=not gpu_tail)
        if gpu_tail:
            _, _, _, ph_at_max, signal, _ = field_result_device(beam_tx, beam_rx, want_arrays=False)
            return ph_at_max, np.angle(signal)
    else:
        beam_tx = Beam(x0=tx[0], y0=tx[1], kx0=0, ky0=1, n_field=nfield)
    _, _, _, ph_at_max, signal, *_ = field_result(beam_tx, beam_rx, device=bool(device))
    return ph_at_max, np.angle(signal)


def phases(den, x, y, *, ant_tx=None, ant_rx=None, a_set=None, nst=-1, nray=251, device=True):
    """Interferometer phase of every TX/RX antenna pair for the density ``den * a`` with ``a`` in
    ``a_set`` (or ``nst + 1`` values from 0 to 1): ``(ph_by_max, ph_by_sig)``, each
    ``[len(a_set), n_antennas]`` (ft2.py:25-71).  Per density scale: one field upload and one
    kernel launch per TX beam instead of ``nray`` Python-level ``Ray`` constructions."""
    Beam._nray = nray                                   # the reference sets the class default too (ft2.py:26)
    if ant_tx is None:                                  # transmitters on the lower half of the antenna circle
        ant_tx = ants() * np.array([1., -1.])
    if ant_rx is None:                                  # receivers mirrored to the upper half
        ant_rx = ant_tx * np.array([1., -1.])
    scales = np.linspace(0, 1, nst + 1) if nst > 0 else np.atleast_1d(a_set)
    rows = []
    for scale in scales:
        nfield = Nfield(den * scale, x=x, y=y)
        rows.append([_pair_phases(nfield, tx, rx, device) for tx, rx in zip(ant_tx, ant_rx)])
        nfield.close()
    out = np.array(rows, float)                         # [scale, antenna, (ph_max, ph_sig)]
    return out[:, :, 0], out[:, :, 1]
