cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_plic_group|k_bound_wave|k_bound_drain|k_un0_worklist|k_face_flux|k_ls_normals|k_mark_near' -s 21 -c 11 -o gpurun_out/prof_r1c python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
