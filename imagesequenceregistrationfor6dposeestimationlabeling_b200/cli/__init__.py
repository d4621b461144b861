"""Drop-in command-line front ends with the reference scripts' flags and file formats
(SURVEY.md section 8(f) row 2):

  python -m <pkg>.cli.verfication --objid 15                       (verfication.py)
  python -m <pkg>.cli.choose_pose --rel_poses 1 --cal_pred 1 ...   (choosePose.py)
  python -m <pkg>.cli.icp --dataset ruapc --objid 1                (icp.py)

Each reads the files the reference script reads, relative to --root (default: the
current directory, like the reference), runs the hot path on the GPU and prints what the
script prints.  `main(argv)` returns the numbers for tests.
"""
