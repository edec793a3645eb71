"""INTEGRATION.md §2 against the real reference tree (build container only: skipped where /root/reference is absent):
after `nlvsgg_b200.lib.install_aliases()` the import lines of the reference's own tools/*.py resolve to the drop-ins, the
untouched reference modules they are used with (lib.AdamW, lib.utils.check_valid_iter) still import from the reference, and
every public callable keeps the reference's parameter names."""
import os
import subprocess
import sys
import textwrap

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lib")), reason="reference tree not present")

SCRIPT = textwrap.dedent('''
    import ast, inspect, sys, types
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float
    sys.path.insert(0, %(repo)r)
    sys.path.append(%(ref)r)
    import nlvsgg_b200.lib as nlv
    nlv.install_aliases()
    # 1. every `from lib.X import ...` / roi_layers line of the four driver scripts that names an aliased module resolves to a drop-in
    checked = 0
    for script in ("train_STTran.py", "test_STTran.py", "train_DSG_DETR.py", "test_DSG_DETR.py"):
        tree = ast.parse(open(%(ref)r + "/tools/" + script).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.ImportFrom) and node.module in nlv.ALIASES:
                mod = __import__(node.module, fromlist=["x"])
                assert mod.__name__ == nlv.ALIASES[node.module], (script, node.module, mod.__name__)
                for a in node.names:
                    if a.name != "*":
                        assert hasattr(mod, a.name), (script, node.module, a.name)
                    else:
                        assert hasattr(mod, "HungarianMatcher")
                    checked += 1
    assert checked >= 8, checked
    # 2. the reference modules used WITH the drop-ins still come from the reference tree
    from lib.AdamW import AdamW
    from lib.utils import check_valid_iter
    assert AdamW.__module__ == "lib.AdamW" and "reference" in inspect.getsourcefile(AdamW)
    assert check_valid_iter.__module__ == "lib.utils"
    # 3. same parameter names as the reference's own definitions (parsed from its sources; nothing of it is executed)
    def ref_params(path, cls, fn):
        tree = ast.parse(open(%(ref)r + "/" + path).read())
        for node in tree.body:
            if cls is None and isinstance(node, ast.FunctionDef) and node.name == fn:
                return [a.arg for a in node.args.args]
            if isinstance(node, ast.ClassDef) and node.name == cls:
                for sub in node.body:
                    if isinstance(sub, ast.FunctionDef) and sub.name == fn:
                        return [a.arg for a in sub.args.args]
        raise KeyError((path, cls, fn))
    def ours(obj):
        return list(inspect.signature(obj).parameters)
    from lib.sttran import STTran
    from lib.dsg_detr import STTran as DSG
    from lib.evaluation_recall import SceneGraphEvaluator
    from lib.track import get_sequence
    from lib.matcher import HungarianMatcher
    from lib.transformer import transformer
    from fasterRCNN.lib.model.roi_layers import ROIAlign, nms
    want = ref_params("lib/sttran.py", "STTran", "__init__")
    assert ours(STTran.__init__)[:len(want)] == want, (ours(STTran.__init__), want)
    want = ref_params("lib/dsg_detr.py", "STTran", "__init__")
    assert ours(DSG.__init__)[:len(want)] == want, (ours(DSG.__init__), want)
    want = ref_params("lib/evaluation_recall.py", "SceneGraphEvaluator", "__init__")
    assert ours(SceneGraphEvaluator.__init__)[:len(want)] == want, (ours(SceneGraphEvaluator.__init__), want)
    ev = [n for n in ast.parse(open(%(ref)r + "/lib/evaluation_recall.py").read()).body if isinstance(n, ast.ClassDef) and n.name == "SceneGraphEvaluator"][0]
    for sub in ev.body:                                   # every public method of the reference class exists, same parameters
        if isinstance(sub, ast.FunctionDef) and not sub.name.startswith("_"):
            assert hasattr(SceneGraphEvaluator, sub.name), sub.name
            assert ours(getattr(SceneGraphEvaluator, sub.name)) == [a.arg for a in sub.args.args], sub.name
    assert ours(get_sequence) == ref_params("lib/track.py", None, "get_sequence")
    assert ours(HungarianMatcher.__init__) == ref_params("lib/matcher.py", "HungarianMatcher", "__init__")
    assert ours(HungarianMatcher.forward) == ref_params("lib/matcher.py", "HungarianMatcher", "forward")
    want = ref_params("lib/transformer.py", "transformer", "__init__")
    assert ours(transformer.__init__)[:len(want)] == want
    assert ours(STTran.forward) == ["self", "entry"] and ours(DSG.forward) == ["self", "entry"]
    # 4. the constructor call of tools/train_STTran.py:79-88 builds a model whose state_dict names are the reference's
    classes = ["__background__"] + open(%(ref)r + "/datasets/AG/object_classes.txt").read().split()
    m = STTran(mode="sgdet", attention_class_num=3, spatial_class_num=6, contact_class_num=17, obj_classes=classes, enc_layer_num=1,
               dec_layer_num=3, transformer_mode="wk", is_wks=True, feat_dim=2048, conf=None)
    from tests import golden_util as G
    assert set(m.state_dict().keys()) == set(G.sttran_template().keys())
    opt = AdamW(m.parameters(), lr=1e-5)           # the reference optimizer accepts the drop-in's parameters
    assert len(opt.param_groups[0]["params"]) == len(list(m.parameters()))
    print("INTEGRATION OK", checked)
''')


def test_reference_tool_imports_resolve_to_the_dropins():
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"repo": repo, "ref": REF}], capture_output=True, text=True, timeout=600, cwd=repo)
    assert r.returncode == 0 and "INTEGRATION OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
