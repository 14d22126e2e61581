/** @file smtx.hxx  The .smtx reader is off the hot path; placeholder so the
 *  reference's include list resolves. */
#pragma once
