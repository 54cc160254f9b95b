"""B200-native backend for CP2K's GPW/GAPW real-space grid hot path
(collocate / integrate task lists).  See DESIGN.md."""
from .grid_api import (  # noqa: F401
    ALL_GRID_FUNCS, BasisSet, GridLayout, GridLibrary, OffloadBuffer, TaskList, load_b200,
    GRID_FUNC_AB, GRID_FUNC_DADB, GRID_BACKEND_B200, GRID_BACKEND_CPU, GRID_BACKEND_REF,
)
