"""River-network / parameter / option containers for the routing path.

These mirror the reference's *inputs* to the routing path, not its internal structures:

* ``RiverNetwork``  -- the variables `read_streamSeg.f90:getData` (:44) pulls from the river-network
  netCDF: ``segId``, ``downSegId``, ``length``, ``slope`` on the ``seg`` dimension and ``hruId``,
  ``hruSegId``, ``area`` on the ``hru`` dimension (+ optional ``width``, ``man_n``, lake variables).
* ``RouteParams``   -- namelist ``param_nml`` groups ``&HSLOPE / &IRF_UH / &KWT`` (`read_param.f90:26-38`,
  defaults `route/ancillary_data/param.nml.default`).
* ``RouteOptions``  -- the control-file keys the routing path reads (`read_control.f90`,
  `public_var.f90:100-145`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

# routing-method ids, public_var.f90:74-80
ACCUM_RUNOFF = 0
IMPULSE_RESPONSE_FUNC = 1
KINEMATIC_WAVE_TRACKING = 2

# flux field ids for get_flux (shared with include/mizuroute_b200.h)
FIELD_REACH_Q = 0
FIELD_REACH_VOL1 = 1
FIELD_REACH_INFLOW = 2
FIELD_WB = 3
FIELD_BASIN_QI = 4
FIELD_BASIN_QR1 = 5
FIELD_BASIN_QR0 = 6
FIELD_REACH_VOL0 = 7


@dataclass
class RouteParams:
    fshape: float = 2.5      # &HSLOPE
    tscale: float = 86400.0
    velo: float = 1.5        # &IRF_UH
    diff: float = 5000.0
    mann_n: float = 0.01     # &KWT
    wscale: float = 0.001


@dataclass
class RouteOptions:
    dt: float = 86400.0                 # <dt_qsim>
    route_opt: str = "1"                # <route_opt> digit string, e.g. "12"
    doesBasinRoute: int = 1
    hw_drain_point: int = 2
    min_length_route: float = 0.0
    is_lake_sim: bool = False
    lakeRegulate: bool = True
    LakeInputOption: int = 0
    runoffMin: float = 0.0
    units_qsim: str = "mm/s"
    floodplain: bool = False            # <floodplain>: finite bankfull depth for the Euler schemes (KW / MC / DW)
    sim_start: Optional[tuple] = None   # (year, month, day, seconds of day) of <sim_start>: calendar of the HYPE lake model
    calendar: str = "standard"          # "standard" (= gregorian, proleptic_gregorian) or "noleap"

    def conv(self):
        """(time_conv, length_conv) exactly as read_control.f90:443-474 derives them."""
        if "/" not in self.units_qsim:
            raise ValueError('expect the character "/" exists in the units string')
        clen, ctime = self.units_qsim.split("/", 1)
        length_conv = {"m": 1.0, "mm": 1.0 / 1000.0}[clen.strip()]
        t = ctime.strip()
        if t in ("d", "day"):
            time_conv = 1.0 / 86400.0
        elif t in ("h", "hr", "hour"):
            time_conv = 1.0 / 3600.0
        elif t in ("s", "sec", "second"):
            time_conv = 1.0
        else:
            raise ValueError("expect the time units of runoff to be day, hour or second")
        return time_conv, length_conv


@dataclass
class RiverNetwork:
    segId: np.ndarray          # int32 [nRch]
    downSegId: np.ndarray      # int32 [nRch]  (<=0: outlet)
    length: np.ndarray         # f64 [nRch]  m
    slope: np.ndarray          # f64 [nRch]  -
    hruId: np.ndarray          # int32 [nHRU]
    hruSegId: np.ndarray       # int32 [nHRU] id of the reach the HRU drains to
    area: np.ndarray           # f64 [nHRU]  m2
    width: Optional[np.ndarray] = None     # f64 [nRch], else wscale*sqrt(totalArea)
    man_n: Optional[np.ndarray] = None     # f64 [nRch], else mann_n
    islake: Optional[np.ndarray] = None    # int32 [nRch]
    lakeModelType: Optional[np.ndarray] = None
    D03_MaxStorage: Optional[np.ndarray] = None
    D03_Coefficient: Optional[np.ndarray] = None
    D03_Power: Optional[np.ndarray] = None
    D03_S0: Optional[np.ndarray] = None
    lake_params: dict = field(default_factory=dict)      # further per-reach lake parameters by reference name, e.g. "HYP_E_emr" (dataTypes.f90:202-254)
    meta: dict = field(default_factory=dict)

    def __post_init__(self):
        i32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        self.segId, self.downSegId = i32(self.segId), i32(self.downSegId)
        self.hruId, self.hruSegId = i32(self.hruId), i32(self.hruSegId)
        self.length, self.slope, self.area = f64(self.length), f64(self.slope), f64(self.area)
        self.width, self.man_n = f64(self.width), f64(self.man_n)
        self.islake, self.lakeModelType = i32(self.islake), i32(self.lakeModelType)
        self.D03_MaxStorage, self.D03_Coefficient = f64(self.D03_MaxStorage), f64(self.D03_Coefficient)
        self.D03_Power, self.D03_S0 = f64(self.D03_Power), f64(self.D03_S0)

    @property
    def nRch(self) -> int:
        return int(self.segId.shape[0])

    @property
    def nHRU(self) -> int:
        return int(self.hruId.shape[0])
