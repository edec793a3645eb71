"""Drop-in mirrors of the reference's lib/ modules on the hot path (same names, constructors, forward(entry)
contract and state_dict keys), backed by the sm_100a kernels.  Put this package's parent on sys.path ahead of the
reference tree — see INTEGRATION.md."""


ALIASES = {
    "lib.sttran": "nlvsgg_b200.lib.sttran", "lib.dsg_detr": "nlvsgg_b200.lib.dsg_detr",
    "lib.transformer": "nlvsgg_b200.lib.transformer", "lib.transformer_wk": "nlvsgg_b200.lib.transformer_wk",
    "lib.evaluation_recall": "nlvsgg_b200.lib.evaluation_recall", "lib.track": "nlvsgg_b200.lib.track",
    "lib.matcher": "nlvsgg_b200.lib.matcher",
    "fasterRCNN.lib.model.roi_layers": "nlvsgg_b200.lib.roi_layers",
    "lib.draw_rectangles.draw_rectangles": "nlvsgg_b200.lib.draw_rectangles.draw_rectangles",
    "lib.fpn.box_intersections_cpu.bbox": "nlvsgg_b200.lib.fpn.box_intersections_cpu.bbox",
}


def install_aliases():
    """INTEGRATION.md §2 in one call: after this, the reference's own `from lib.sttran import STTran`,
    `from lib.evaluation_recall import SceneGraphEvaluator`, `from lib.track import get_sequence`, `from lib.matcher import *`,
    `from fasterRCNN.lib.model.roi_layers import ROIAlign, nms` ... resolve to the sm_100a drop-ins; everything else of the
    reference tree (lib.AdamW, lib.utils, lib.config, the dataloader) is untouched.  Call it before those imports."""
    import importlib
    import sys
    import types
    for ref_name, ours in ALIASES.items():
        sys.modules[ref_name] = importlib.import_module(ours)
    for pkg in ("fasterRCNN", "fasterRCNN.lib", "fasterRCNN.lib.model"):      # the vendored detector package need not be built
        if pkg not in sys.modules:
            sys.modules[pkg] = types.ModuleType(pkg)
            sys.modules[pkg].__path__ = []
