"""KITTI directory layout (reference src/e2eflow/kitti/data.py:89-153, core/data.py).

The reference downloads missing archives (KITTI raw sequences, flow 2012 / 2015) on construction;
this environment has no network and downloading is outside the hot path, so a missing directory
is an error that names what is expected."""
import os


class KITTIData():
    dirs = ['data_stereo_flow', 'data_scene_flow', 'kitti_raw']

    def __init__(self, data_dir, stat_log_dir=None, development=True, fast_dir=None, require=('kitti_raw',)):
        self.development = development
        self.data_dir = data_dir
        self.stat_log_dir = stat_log_dir
        self.fast_dir = fast_dir
        self.current_dir = fast_dir or data_dir
        for d in require:
            if not os.path.isdir(os.path.join(self.current_dir, d)):
                raise FileNotFoundError(
                    "%s not found under %s (expected the layout the reference's downloader creates: "
                    "kitti_raw/<date>/<date>_drive_<n>_extract/image_0{2,3}/data/*.png, "
                    "data_stereo_flow/, data_scene_flow/)" % (d, self.current_dir))

    def get_raw_dirs(self):
        top_dir = os.path.join(self.current_dir, 'kitti_raw')
        dirs = []
        for date in os.listdir(top_dir):
            date_path = os.path.join(top_dir, date)
            for extract in os.listdir(date_path):
                extract_path = os.path.join(date_path, extract)
                dirs.extend([os.path.join(extract_path, 'image_02/data'),
                             os.path.join(extract_path, 'image_03/data')])
        return dirs

    def get_raw_files(self):
        return [os.path.join(d, p) for d in self.get_raw_dirs() for p in os.listdir(d)]
