"""Patch the B200 hot path into an imported copy of the reference's `src` package (single process).
For the spawned tracker/mapper processes use the one-line edits in INTEGRATION.md instead: `spawn`
re-imports `src` in the children and module patches made in the parent are not inherited.
"""
from __future__ import annotations

import importlib


def install(src_package: str = "src") -> None:
    from . import decoders, mapper, renderer, tracker

    dec_mod = importlib.import_module(f"{src_package}.networks.decoders")
    dec_mod.Decoders = decoders.Decoders
    for name in (f"{src_package}.networks.config", f"{src_package}.networks"):
        try:
            m = importlib.import_module(name)
            if hasattr(m, "Decoders"):
                m.Decoders = decoders.Decoders
        except ImportError:
            pass
    rnd_mod = importlib.import_module(f"{src_package}.utils.Renderer")
    rnd_mod.Renderer = renderer.Renderer
    try:
        es = importlib.import_module(f"{src_package}.ESLAM")
        es.Renderer = renderer.Renderer
    except ImportError:
        pass
    trk_mod = importlib.import_module(f"{src_package}.Tracker")
    trk_mod.Tracker.optimize_tracking = tracker.optimize_tracking
    trk_mod.Tracker.track_frame = tracker.track_frame
    map_mod = importlib.import_module(f"{src_package}.Mapper")
    map_mod.Mapper.optimize_mapping = mapper.optimize_mapping
    map_mod.Mapper.keyframe_selection_overlap = mapper.keyframe_selection_overlap
