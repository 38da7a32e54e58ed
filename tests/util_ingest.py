"""timeMap_sim_forc (get_basin_runoff.f90:256-369) restated for the tests of the device forcing ingest: which forcing records
lie under every simulation step and what share of the step each covers."""
import numpy as np


def time_map(n_steps, dt, n_rec, dt_ro, t_offset=0.0, tol=1e-9):
    """CSR (rec_ptr, rec_idx, rec_frac); a step inside one record gets that record with share 1."""
    ptr, idx, frac = [0], [], []
    for k in range(n_steps):
        sim1 = t_offset + k * dt
        sim2 = sim1 + dt
        front = max(int(np.floor(sim1 / dt_ro + tol)), 0)
        end = min(max(int(np.ceil(sim2 / dt_ro - tol)) - 1, 0), n_rec - 1)
        assert front <= end
        for r in range(front, end + 1):
            idx.append(r)
            if front == end:
                frac.append(1.0)
            elif r == front:
                frac.append(((r + 1) * dt_ro - sim1) / dt)
            elif r == end:
                frac.append((sim2 - r * dt_ro) / dt)
            else:
                frac.append(dt_ro / dt)
        ptr.append(len(idx))
    return np.array(ptr, dtype=np.int32), np.array(idx, dtype=np.int32), np.array(frac, dtype=np.float64)
