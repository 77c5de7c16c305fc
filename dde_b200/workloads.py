"""Synthetic workloads of BASELINE.json.

Inputs are generated once on the host with a fixed generator and the SAME
arrays are fed to the GPU path and to the CPU checker, so no cross-
implementation RNG matching is needed.  Generator: splitmix64 seeded with
0x9E3779B97F4A7C15, u = (z >> 11) * 2^-53, draws in trajectory-major order.
"""
import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def splitmix64_uniform(count, start=0):
    """`count` uniforms in [0, 1): draws start .. start + count - 1 of the stream."""
    with np.errstate(over="ignore"):
        idx = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = _GOLDEN + idx * _GOLDEN  # state after k+1 increments of the seed
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def mackey_glass(n_traj, first=0):
    """C2: beta = 0.15 + 0.10u, gamma = 0.08 + 0.04u, n = 8 + 4u, y0 = 0.5 + u
    (4 draws per trajectory, in that order); times 0, 50, ..., 500; dop853,
    rtol = atol = 1e-6, n_history = 64.  `first` offsets into the stream so a
    shard is the same as the slice of the full batch."""
    u = splitmix64_uniform(4 * n_traj, start=4 * first).reshape(n_traj, 4)
    parms = np.stack([0.15 + 0.10 * u[:, 0], 0.08 + 0.04 * u[:, 1], 8.0 + 4.0 * u[:, 2]], axis=1)
    y0 = (0.5 + u[:, 3])[:, None]
    times = np.linspace(0.0, 500.0, 11)
    kwargs = dict(method="dopri853", rtol=1e-6, atol=1e-6, n_history=64)
    return "mackey_glass", np.ascontiguousarray(y0), np.ascontiguousarray(parms), times, kwargs


def lorenz_sweep(n_traj, first=0, n_times=1001, t_end=10.0):
    """C3 (Lorenz half): sigma = 8 + 4u, rho = 20 + 16u, b = 2 + 1.5u, y0 = (10, 1, 1);
    1001 times on [0, 10], return_initial = False; dopri5, tol 1e-6."""
    u = splitmix64_uniform(3 * n_traj, start=3 * first).reshape(n_traj, 3)
    parms = np.stack([8.0 + 4.0 * u[:, 0], 20.0 + 16.0 * u[:, 1], 2.0 + 1.5 * u[:, 2]], axis=1)
    y0 = np.tile(np.array([10.0, 1.0, 1.0]), (n_traj, 1))
    times = np.linspace(0.0, t_end, n_times)
    kwargs = dict(method="dopri5", rtol=1e-6, atol=1e-6, n_history=0, return_initial=False)
    return "lorenz", y0, np.ascontiguousarray(parms), times, kwargs


def rossler_sweep(n_traj, first=0, n_times=1001, t_end=10.0):
    """C3 (Roessler half): a = 0.1 + 0.2u, b = 0.1 + 0.2u, c = 4 + 4u, y0 = (1, 1, 1)."""
    u = splitmix64_uniform(3 * n_traj, start=3 * first).reshape(n_traj, 3)
    parms = np.stack([0.1 + 0.2 * u[:, 0], 0.1 + 0.2 * u[:, 1], 4.0 + 4.0 * u[:, 2]], axis=1)
    y0 = np.ones((n_traj, 3))
    times = np.linspace(0.0, t_end, n_times)
    kwargs = dict(method="dopri5", rtol=1e-6, atol=1e-6, n_history=0, return_initial=False)
    return "rossler", y0, np.ascontiguousarray(parms), times, kwargs


def sir_delay(n_rep, first=0, n_steps=10000):
    """C4: beta = 0.2 + 0.2u, gamma = 0.05 + 0.1u, y0 = (0.99, 0.01, 0);
    steps = (0, n_steps), return_initial = False, n_history = 16."""
    u = splitmix64_uniform(2 * n_rep, start=2 * first).reshape(n_rep, 2)
    parms = np.stack([0.2 + 0.2 * u[:, 0], 0.05 + 0.1 * u[:, 1]], axis=1)
    y0 = np.tile(np.array([0.99, 0.01, 0.0]), (n_rep, 1))
    steps = np.array([0, n_steps], dtype=np.uint64)
    kwargs = dict(n_history=16, return_initial=False)
    return "sir_delay", y0, np.ascontiguousarray(parms), steps, kwargs


def seir_age(n_traj, first=0, t_end=200.0, n_times=21):
    """C5: age-structured SEIR, 128 age classes x (S, E, I, R) = 512 states, lags 14 and 21;
    beta = 8 + 4u, sigma = 1 / (2 + 2u), delta = 1 / (14 + 14u); y0: the population of 1e7
    spread evenly over the age classes, all susceptible but one infective in class 0;
    dop853, rtol = atol = 1e-6, n_history = 128, times 0 .. t_end."""
    u = splitmix64_uniform(3 * n_traj, start=3 * first).reshape(n_traj, 3)
    parms = np.stack([8.0 + 4.0 * u[:, 0], 1.0 / (2.0 + 2.0 * u[:, 1]),
                      1.0 / (14.0 + 14.0 * u[:, 2])], axis=1)
    A = 128
    y0 = np.zeros((n_traj, 4 * A))
    y0[:, :A] = 1e7 / A
    y0[:, 0] -= 1.0
    y0[:, 2 * A] = 1.0
    times = np.linspace(0.0, t_end, n_times)
    kwargs = dict(method="dopri853", rtol=1e-6, atol=1e-6, n_history=128)
    return "seir_age", y0, np.ascontiguousarray(parms), times, kwargs
