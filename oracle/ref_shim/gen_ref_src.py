#!/usr/bin/env python3
"""ORACLE BUILD INFRASTRUCTURE (not product code).

Reads the reference's kernel TEXT files from <ref>/src/kernel (they are templates with
_placeholders_, runtime-compiled by jitify in the reference: src/core/utils/JitHelper.cpp:50-147,
src/DEM/APIPrivate.cpp:294-329) and performs the same placeholder substitution the reference performs
at Initialize() (equip* functions, src/DEM/APIPrivate.cpp:1381-2133, src/DEM/Models.h:254-379), except
that the "jitified constant tables" are bound to runtime pointers owned by oracle/ref_shim/ref_harness.cpp
instead of array literals.  Output goes to a scratch directory given on the command line; nothing of the
reference's text is ever written into the repository.

usage: gen_ref_src.py <reference_root> <out_dir>
"""
import os
import re
import sys


def read(ref, rel):
    with open(os.path.join(ref, "src", "kernel", rel), "r") as f:
        return f.read()


def subst(text, mapping):
    # longest keys first so "_forceModelIngredientAcqForA_" is not clobbered by shorter keys
    for k in sorted(mapping, key=len, reverse=True):
        text = text.replace(k, mapping[k])
    return text


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    pol = "DEMCustomizablePolicies/"

    common = {
        "_kernelIncludes_;": "",
        "_clumpTemplateDefs_;": "",
        "_analyticalEntityDefs_;": "",
        "_materialDefs_;": "",
        "_massDefs_;": "",
        "_moiDefs_;": "",
        "_forceModelPrerequisites_;": "",
        "_nvXp2_": "g_nvXp2",
        "_nvYp2_": "g_nvYp2",
        "_voxelSize_": "g_voxelSize",
        "_l_": "g_l",
        # jitified template / mass acquisition (the reference's defaults, API.h:1397-1399)
        "_componentAcqStrat_": read(ref, pol + "ClumpCompAcqStratAllJitify.cu"),
        "_massAcqStrat_": read(ref, pol + "MassAcqStratJitify.cu"),
        "_moiAcqStrat_": read(ref, pol + "MOIAcqStratJitify.cu"),
    }

    # ---- force kernel: what equipForceModel + equip_force_model_ingr_acq + equip_contact_wildcards emit ----
    ingr_def = ("float ts = simParams->h;\n"
                "deme::family_t AOwnerFamily;\n"
                "deme::family_t BOwnerFamily;\n"
                "float3 ALinVel, BLinVel;\n"
                "float3 ARotVel, BRotVel;\n")
    acq = ("{S}OwnerFamily = granData->familyID[myOwner];\n"
           "{S}LinVel.x = granData->vX[myOwner];\n{S}LinVel.y = granData->vY[myOwner];\n"
           "{S}LinVel.z = granData->vZ[myOwner];\n"
           "{S}RotVel.x = granData->omgBarX[myOwner];\n{S}RotVel.y = granData->omgBarY[myOwner];\n"
           "{S}RotVel.z = granData->omgBarZ[myOwner];\n")
    wc_names = ["delta_tan_x", "delta_tan_y", "delta_tan_z", "delta_time"]  # std::set order, AuxClasses.cpp:761
    wc_acq = "".join("float %s = granData->contactWildcards[%d][myContactID];\n" % (n, i)
                     for i, n in enumerate(wc_names))
    wc_wb = "".join("granData->contactWildcards[%d][myContactID] = %s;\n" % (i, n) for i, n in enumerate(wc_names))
    wc_destroy = "".join("%s = 0;\n" % n for n in wc_names)

    force_tpl = read(ref, "DEMCalcForceKernels.cu")
    for tag, model_file, hist in (("full", "FullHertzianForceModel.cu", True),
                                  ("frictionless", "FrictionlessHertzianForceModel.cu", False)):
        m = dict(common)
        m.update({
            "_DEMForceModel_": read(ref, pol + model_file),
            "_forceModelIngredientDefinition_": ingr_def,
            "_forceModelIngredientAcqForA_": acq.format(S="A"),
            "_forceModelIngredientAcqForB_": acq.format(S="B"),
            "_forceModelGeoWildcardAcqForSph_": " ",
            "_forceModelGeoWildcardAcqForTri_": " ",
            "_forceModelGeoWildcardAcqForAnal_": " ",
            "_forceModelOwnerWildcardWrite_": " ",
            "_forceModelContactWildcardAcq_": wc_acq if hist else " ",
            "_forceModelContactWildcardWrite_": wc_wb if hist else " ",
            "_forceModelContactWildcardDestroy_": wc_destroy if hist else " ",
            "_forceCollectInPlaceStrat_": " ",
            "_contactInfoWrite_": read(ref, pol + "ContactInfoWriteBack.cu"),
        })
        with open(os.path.join(out, "calcforce_%s.inc" % tag), "w") as f:
            f.write(subst(force_tpl, m))

    # ---- force -> acceleration (atomics path, the reference default) ----
    txt = subst(read(ref, "DEMCollectForceKernels_Compact.cu"), common)
    # the file carries its own literal objOwner[] table; the harness binds objOwner to a runtime pointer
    txt = re.sub(r"__constant__ __device__ deme::bodyID_t objOwner\[\] = \{_objOwner_\};", "", txt)
    with open(os.path.join(out, "collect_compact.inc"), "w") as f:
        f.write(txt)

    with open(os.path.join(out, "prepforce.inc"), "w") as f:
        f.write(subst(read(ref, "DEMPrepForceKernels.cu"), common))
    with open(os.path.join(out, "misc.inc"), "w") as f:
        f.write(subst(read(ref, "DEMMiscKernels.cu"), common))
    with open(os.path.join(out, "binsphere.inc"), "w") as f:
        f.write(subst(read(ref, "DEMBinSphereKernels.cu"), common))
    with open(os.path.join(out, "history.inc"), "w") as f:
        f.write(subst(read(ref, "DEMHistoryMappingKernels.cu"), common))
    with open(os.path.join(out, "contact_ss.inc"), "w") as f:
        f.write(subst(read(ref, "DEMContactKernels_SphereSphere.cu"), common))
    # ---- sphere--triangle broad phase: facet sandwich + facet -> bin registration (per-thread kernels, executed), and the
    #      per-bin sweep (block-cooperative: only its per-thread helper functions are called) ----
    with open(os.path.join(out, "bintriangle.inc"), "w") as f:
        f.write(subst(read(ref, "DEMBinTriangleKernels.cu"), common))
    with open(os.path.join(out, "contact_st.inc"), "w") as f:
        f.write(subst(read(ref, "DEMContactKernels_SphereTriangle.cu"), common))

    # ---- integration: family prescriptions are table driven (constants only), see
    #      equipFamilyPrescribedMotions, APIPrivate.cpp:1601-1708 ----
    vel = ("case 0 ... 255: { const OrcPrescription& P_ = g_presc[family]; if (P_.used) {"
           "{ if (P_.hasLinVel[0]) vX = P_.linVel[0]; if (P_.hasLinVel[1]) vY = P_.linVel[1];"
           " if (P_.hasLinVel[2]) vZ = P_.linVel[2]; }"
           "{ if (P_.hasRotVel[0]) omgBarX = P_.rotVel[0]; if (P_.hasRotVel[1]) omgBarY = P_.rotVel[1];"
           " if (P_.hasRotVel[2]) omgBarZ = P_.rotVel[2]; }"
           "LinVelXPrescribed = P_.linVelPrescribed[0]; LinVelYPrescribed = P_.linVelPrescribed[1];"
           "LinVelZPrescribed = P_.linVelPrescribed[2]; RotVelXPrescribed = P_.rotVelPrescribed[0];"
           "RotVelYPrescribed = P_.rotVelPrescribed[1]; RotVelZPrescribed = P_.rotVelPrescribed[2]; } break; }")
    pos = ("case 0 ... 255: { const OrcPrescription& P_ = g_presc[family]; if (P_.used) {"
           "{ if (P_.hasLinPos[0]) X = P_.linPos[0]; if (P_.hasLinPos[1]) Y = P_.linPos[1];"
           " if (P_.hasLinPos[2]) Z = P_.linPos[2]; }"
           "LinXPrescribed = P_.linPosPrescribed[0]; LinYPrescribed = P_.linPosPrescribed[1];"
           "LinZPrescribed = P_.linPosPrescribed[2]; RotPrescribed = P_.rotPosPrescribed; } break; }")
    acc = ("case 0 ... 255: { const OrcPrescription& P_ = g_presc[family]; if (P_.used) {"
           "{ if (P_.hasAcc[0]) accX = P_.acc[0]; if (P_.hasAcc[1]) accY = P_.acc[1];"
           " if (P_.hasAcc[2]) accZ = P_.acc[2]; }"
           "{ if (P_.hasAngAcc[0]) angAccX = P_.angAcc[0]; if (P_.hasAngAcc[1]) angAccY = P_.angAcc[1];"
           " if (P_.hasAngAcc[2]) angAccZ = P_.angAcc[2]; } } break; }")
    integ_tpl = read(ref, "DEMIntegrationKernels.cu")
    for tag, fname in (("euler", "IntegrationVelPassOnForwardEuler.cu"),
                       ("centered", "IntegrationVelPassOnCenteredDiff.cu"),
                       ("taylor", "IntegrationVelPassOnExtendedTaylor.cu")):
        m = dict(common)
        m.update({
            "_velPrescriptionStrategy_": vel,
            "_posPrescriptionStrategy_": pos,
            "_accPrescriptionStrategy_": acc,
            "_integrationVelocityPassOnStrategy_": read(ref, pol + fname),
        })
        with open(os.path.join(out, "integrate_%s.inc" % tag), "w") as f:
            f.write(subst(integ_tpl, m))


if __name__ == "__main__":
    main()
