"""Host-side mirrors of the reference's entry points: same function names,
argument dicts, return shapes and error behaviour as the module-level functions the
reference's GUI classes call, with the numeric work routed to libipb200.so.

  Fluor_INT               _process_key_task(task)            src/INT/Fluor_INT.py:795
  FA_Analyzer             analyze_fa_crop(...), batch body   src/INT/FA_Analyzer.py:123,939
  fret_ratio_builder      process_one_stage(...)             src/FRET/fret_ratio_builder.py:429
  Nesprin2_FRET_Builder   run_pipeline(p)                    src/FRET/Nesprin2_FRET_Builder.py:1331
  MOR_by_ROI              morphology_from_polygon(...)       src/MOR_by_ROI.py:211
  roi_channel_cropper     run_crop numeric body              src/roi_channel_cropper.py:778-969

The Tk GUIs, matplotlib figure rendering (PNG overlays) and the ROI drawers stay on the host
and are out of scope; every mirror has a headless `run_*` driver that
takes the same parameter dict the GUI assembles.  Input images must be 8/16-bit integer TIFFs
(what the microscopes write); there is no CPU fallback for the numeric path.
"""
