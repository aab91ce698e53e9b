# developer: correctness subset + timing of the working build
python -m pytest tests/test_gpu_planner.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -4
for d in dense_cluster_on_path friendly_ring survey_c2 clutter_in_reach pillars_in_reach all_ties_far_obstacles empty_cloud; do
  python tools/family_dev.py --replay $d 400 2>&1 | tail -1
done
python tools/sweep_prof_dev.py dense_cluster_on_path 64 2>&1 | tail -1
python tools/sweep_prof_dev.py friendly_ring 64 2>&1 | tail -1
export KOMPASS_B200_LIB=$PWD/kompass-core_b200/lib/variants/libkompass_b200_dbg.so
python tools/stamps_dev.py friendly_ring dense_cluster_on_path survey_c2 clutter_in_reach pillars_in_reach all_ties_far_obstacles 2>&1 | tail -6
python tools/cycle_stamps_dev.py friendly_ring dense_cluster_on_path survey_c2 2>&1 | tail -36
