"""Shadows ``gsplat/examples/simple_trainer.py`` -- the script the reference's Train-Stage1 component launches
(source/container/src/main.py:1339-1347, ``command="gsplat/examples/simple_trainer.py"``, cwd = the container's working
directory): with ``shim/`` as that directory (or on PYTHONPATH) the unchanged component command runs this package's trainer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))

from pipeline_pointcloud_b200.simple_trainer import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
