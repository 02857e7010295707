"""The usage block of README.md is executed as written (needs a B200): documentation that runs."""
import os
import re

import numpy as np
import pytest


@pytest.mark.gpu
def test_readme_usage_block_runs():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "README.md")).read()
    code = re.search(r"```python\n(.*?)```", text, re.S).group(1)
    ns = {"np": np}
    exec(compile(code, "README.md", "exec"), ns)
    assert np.isfinite(np.nansum(ns["img"])) and ns["img"].shape == (2048, 2048)
    assert ns["flux"].sum() == pytest.approx(1.0) and len(ns["ctfs"]) == 150 and len(ns["table"]) == 3
    assert np.nansum(ns["shadow"]) > 0 and ns["prof"].eps[0] > ns["prof"].eps[-1]
