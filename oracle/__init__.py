"""CPU oracle for the correspondence hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (slam_indoor_code_b200 + libslamb200.so) never does.

  oracle.c_oracle   ctypes view of oracle/liboracle.so (C restatement, corr_oracle.c)
  oracle.np_oracle  NumPy restatement of the same functions (independent second statement)
  oracle.synth      seeded synthetic inputs of SURVEY.md section 8(d)
  oracle.gen_golden regenerates tests/golden/*.npz from cv2 (the reference's arithmetic owner)
"""
